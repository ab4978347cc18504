"""CPU-side checks of what the kernels rely on (no GPU needed):

* the SASS of the built library holds the instructions DESIGN.md claims for the hot kernels (bulk copies + mbarriers in
  k_fast, 3-input DPX min/max, TMA tensor-tile loads in the default BRIEF kernel, packed f32x2 in k_blur) - and, the
  regression this file exists for, NO packed FMA in the BRIEF steering: ptxas contracts `mul.rn.f32x2` + `add.rn.f32x2`
  into one `FFMA2` (a single rounding where cv::ORB rounds twice), so the kernels keep those sums scalar;
* the algebra of k_fast's score tree (shared 8-windows) and of the biased-byte float conversion, by brute force in numpy.
"""
import collections
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "orb-slam3-rust_b200", "vslam_b200", "libvslam_b200.so")


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = {}
    for block in re.split(r"\n\s*Function : ", out)[1:]:
        name = block.split("\n", 1)[0].strip()
        ops = collections.Counter()
        for line in block.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Za-z0-9_.]+)", line)
            if m:
                ops[m.group(1).rstrip(";")] += 1
        kernels[name] = ops
    return kernels


def ops_of(sass, mangled_fragment):
    hits = [v for k, v in sass.items() if mangled_fragment in k]
    assert len(hits) == 1, f"{mangled_fragment}: {len(hits)} kernels match"
    return hits[0]


def count(ops, prefix):
    return sum(n for op, n in ops.items() if op.startswith(prefix))


def test_k_fast_stages_with_bulk_copies_and_scores_with_dpx(sass):
    ops = ops_of(sass, "2vk6k_fastE")
    assert count(ops, "UBLKCP") == 2, "the pixel tile is two cp.async.bulk copies"
    assert count(ops, "SYNCS") >= 4, "mbarrier init / expect_tx / try_wait"
    assert count(ops, "VIMNMX3.U16x2") >= 40 and count(ops, "VIMNMX.U16x2") >= 90, "DPX min/max score tree on u16x2 lanes"
    assert count(ops, "VABSDIFF4") >= 56, "byte-SIMD pre-test, seven unrolled row steps"
    assert count(ops, "LDL") == 0 and count(ops, "STL") == 0, "no local-memory traffic (register window indices are constants)"


def test_default_brief_kernel_uses_tma_tile_loads_and_no_packed_fma(sass):
    tma = ops_of(sass, "2vk11k_brief_tmaE")
    assert count(tma, "UTMALDG") >= 2, "blurred + un-blurred patch arrive by cp.async.bulk.tensor"
    for frag in ("2vk11k_brief_tmaE", "2vk7k_briefE"):
        ops = ops_of(sass, frag)
        assert count(ops, "FMUL2") == 32 and count(ops, "FADD2") == 32, "8 tests x (4 products + 4 bias / un-bias adds), packed"
        assert count(ops, "FFMA2") == 0, ("ptxas contracted the steering's packed mul + add into FFMA2: one rounding where the "
                                          "reference rounds twice (DESIGN.md 5, 'a ptxas trap')")


def test_k_blur_uses_packed_f32(sass):
    ops = ops_of(sass, "2vk6k_blurE")
    assert count(ops, "FFMA2") >= 300 and count(ops, "LDGSTS") >= 100
    assert count(ops, "LDL") == 0 and count(ops, "STL") == 0


def test_no_tensor_core_instructions_anywhere(sass):
    for name, ops in sass.items():
        assert count(ops, "HMMA") == 0 and count(ops, "UTCHMMA") == 0 and count(ops, "UTCQMMA") == 0, name


# ------------------------------------------------------------------------------------------------ algebra
def brute_force_fast(p):
    hi = np.full(len(p), -1)
    lo = np.full(len(p), 1 << 20)
    for k in range(16):
        idx = [(k + j) % 16 for j in range(9)]
        hi = np.maximum(hi, p[:, idx].min(1))
        lo = np.minimum(lo, p[:, idx].max(1))
    return hi, lo


def shared_window_tree(p):
    """orb_kernels.cuh fast_score2: the 9-windows starting at k and k+1 share {k+1..k+8}, and
    max(min9_k, min9_{k+1}) = min(min8, max(p_k, p_{k+9}))."""
    P = lambda i: p[:, i % 16]  # noqa: E731
    a2 = [np.minimum(P(2 * i + 1), P(2 * i + 2)) for i in range(8)]
    b2 = [np.maximum(P(2 * i + 1), P(2 * i + 2)) for i in range(8)]
    a4 = [np.minimum(a2[i], a2[(i + 1) % 8]) for i in range(8)]
    b4 = [np.maximum(b2[i], b2[(i + 1) % 8]) for i in range(8)]
    r = [np.minimum(np.minimum(a4[i], a4[(i + 2) % 8]), np.maximum(P(2 * i), P(2 * i + 9))) for i in range(8)]
    q = [np.maximum(np.maximum(b4[i], b4[(i + 2) % 8]), np.minimum(P(2 * i), P(2 * i + 9))) for i in range(8)]
    return np.max(r, axis=0), np.min(q, axis=0)


def test_score_tree_over_shared_windows_equals_the_sixteen_nine_windows():
    rng = np.random.default_rng(7)
    p = rng.integers(0, 256, size=(300_000, 16))
    p[:60_000] = rng.integers(0, 3, size=(60_000, 16)) * 100       # few distinct values: ties everywhere
    p[60_000:90_000] = np.sort(p[60_000:90_000], axis=1)            # monotone circles
    p[90_000:120_000, ::2] = 255                                    # alternating extremes
    hi, lo = brute_force_fast(p)
    hi2, lo2 = shared_window_tree(p)
    assert np.array_equal(hi, hi2) and np.array_equal(lo, lo2)


def test_biased_byte_to_float_is_exact():
    """steered_brief_byte: byte u = v + 128 dropped into the mantissa of 2^23, then + (-(2^23 + 128)) gives v exactly."""
    u = np.arange(256, dtype=np.uint32)
    as_float = (np.uint32(0x4B000000) | u).view(np.float32)
    v = as_float + np.float32(-8388736.0)
    assert np.array_equal(v, (u.astype(np.int64) - 128).astype(np.float32))


def test_segment_quotient_by_reciprocal_is_exact_for_every_supported_width():
    """k_fast: seg = umulhi(li, 0xffffffff / ngx + 1) must equal li // ngx for li < 5 * ngx, ngx = 2 .. 2048."""
    for ngx in range(2, 2049):
        inv = 0xFFFFFFFF // ngx + 1
        li = np.arange(5 * ngx, dtype=np.uint64)
        assert np.array_equal((li * np.uint64(inv)) >> np.uint64(32), li // np.uint64(ngx)), ngx
