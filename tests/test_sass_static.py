"""Static checks on the SASS of the built library (cuobjdump, no GPU):
  * the tensor-core kernels really are tcgen05 / TMEM / bulk-TMA code (UTCHMMA, LDTM / STTM, UBLKCP);
  * the direct convolution stages its tiles with zero-filling cp.async (LDGSTS ... ZFILL);
  * in every tensor-core kernel the epilogue's slot-release `mbarrier.arrive` is issued AFTER all arithmetic that
    consumes the shared-memory / TMEM loads of that slot.  ptxas once hoisted it above those consumers, which let the
    producer's next bulk copy overtake loads in flight (rare corrupted rows, DESIGN.md 4.1 "Slot-release ordering");
    `mbar_arrive_after` pins it - this test fails if a compiler or code change undoes that."""
import re
import shutil
import subprocess

import pytest

from jolideco_b200 import build

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="needs cuobjdump")


@pytest.fixture(scope="module")
def functions():
    lib = build.build()
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out = {}
    for chunk in re.split(r"\n\s*Function : ", txt)[1:]:
        name, body = chunk.split("\n", 1)
        out[name.strip()] = [line for line in body.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line)]
    return out


def test_tensor_core_kernels_use_tcgen05_tmem_and_bulk_tma(functions):
    kernels = {n: b for n, b in functions.items() if re.search(r"gmm_fwd_tc|gmm_bwd_lse_tc|gmm_fwd_tc16", n)}
    # 4 + 4 forward variants (tile / stream-K), 4 FP16, 4 mixed TF32/FP16, 2 x 4 two-tile kernels (both recipes), 1 logsumexp backward
    assert len(kernels) >= 25
    for name, body in kernels.items():
        text = "\n".join(body)
        for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR"):
            assert mnemonic in text, f"{mnemonic} missing in {name}"


def test_direct_convolution_stages_with_zero_filling_cp_async(functions):
    kernels = {n: b for n, b in functions.items() if "conv3_kernel" in n}
    assert len(kernels) == 24  # 2 directions x 3 tile shapes x 4 tap tails
    for name, body in kernels.items():
        text = "\n".join(body)
        assert "LDGSTS" in text and "ZFILL" in text, name


def test_slot_release_arrive_follows_the_consumers_of_the_loads(functions):
    checked = 0
    for name, body in functions.items():
        if not re.search(r"gmm_fwd_tc|gmm_bwd_lse_tc|gmm_fwd_tc16", name):
            continue
        arrives = [i for i, line in enumerate(body)
                   if "SYNCS.ARRIVE.TRANS64.A1T0" in line and "@" in line.split("SYNCS")[0]]
        assert arrives, f"no predicated slot-release arrive found in {name}"
        releases = 0
        for i in arrives:
            loads = [k for k in range(i) if re.search(r"\bLDTM\b|LDTM\.", body[k])]
            if not loads:
                continue  # an arrive of another role (gather -> MMA hand-over of jd_gmm_tcm.cu), not a slot release
            # the slot release closes the epilogue's arithmetic on the accumulator it frees: no branch in between
            between = body[loads[-1] + 1:i]
            if any(re.search(r"\bBAR\b", line) for line in between):
                continue  # end-of-segment arrive (A buffer hand-back), separated from the loads by barriers
            math_before = sum(1 for line in between if re.search(r"FFMA|FADD|FMUL", line))
            math_after = 0
            for k in range(i + 1, min(i + 200, len(body))):
                if re.search(r"\bBRA\b|\bEXIT\b", body[k]):
                    break
                math_after += bool(re.search(r"FFMA|FADD|FMUL", body[k]))
            # (>= 16 packed or >= 40 scalar FMAs of the squares precede it, nothing that consumes a load follows)
            assert math_before >= 20 and math_after == 0, (name, math_before, math_after)
            releases += 1
        assert releases >= 1, name
        checked += releases
    assert checked >= 25


def test_bucket_kernel_gathers_with_asynchronous_copies(functions):
    """gmm_bwd_bucket8_kernel: Lam_k (16-byte) and the patch elements (4-byte) enter shared memory as cp.async (LDGSTS),
    all of them issued before the wait.  With plain loads ptxas interleaved each patch's two loads with that patch's
    shuffle reduction - one L2 round trip per patch, 63 us instead of 33 us (profiles/r02_summary.md)."""
    kernels = {n: b for n, b in functions.items() if "gmm_bwd_bucket8_kernel" in n}
    assert len(kernels) == 1
    body = next(iter(kernels.values()))
    copies = [i for i, line in enumerate(body) if "LDGSTS" in line]
    wide = [i for i in copies if ".128" in body[i]]
    assert len(wide) >= 8 and len(copies) - len(wide) >= 16  # 8 x 16 B per thread of Lam_k, 2 copies x 8 unrolled patches
    waits = [i for i, line in enumerate(body) if "DEPBAR" in line and "LDGDEPBAR" not in line]
    assert waits and waits[0] > copies[-1]  # one wait, after the last copy
    # no warp shuffle (the centring) between the first copy and the wait
    assert not any("SHFL" in line for line in body[copies[0]:waits[0]])
    assert sum("FFMA" in line for line in body) >= 256
