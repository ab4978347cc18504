"""CPU oracle for the stereo front end — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product path (orb-slam3-rust_b200/) never does; it fails loudly when its CUDA library is missing.

Two layers:
  * `cv2_ref`     — the reference's actual extraction engine: cv::ORB::detectAndCompute through Python cv2
                    (the same C++ function /root/reference/src/tracking/frame/stereo.rs:72-73 reaches through the
                    `opencv` crate).  This PINS the extraction restatement.
  * `orb_oracle`  — ctypes bindings of liborb_oracle.so: the stage-by-stage C++ restatement of cv::ORB
                    (orb_oracle.cpp) and of the reference's Rust loops (stereo_oracle.cpp).
"""
