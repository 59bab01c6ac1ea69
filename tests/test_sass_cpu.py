"""The built library really is sm_100a tensor-core code: its SASS must contain the Blackwell mnemonics
(B200_PROFILING.md): UTCHMMA (tcgen05.mma, incl. the .2CTA pair form), UTMALDG (TMA tensor loads), LDTM
(tcgen05.ld from tensor memory), UTCBAR (tcgen05.commit) — and none of the legacy HMMA / wgmma paths."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

LIB = Path(__file__).resolve().parent.parent / "skyrl_b200" / "libskyb200.so"


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not LIB.exists(), reason="needs cuobjdump and the built library")
def test_library_sass_uses_tcgen05_tma_and_tensor_memory():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in sass
    count = lambda pat: len(re.findall(pat, sass))
    assert count(r"\bUTCHMMA\.2CTA\b") > 0, "no cta_group::2 tcgen05.mma"
    assert count(r"\bUTMALDG\.2D") > 0, "no TMA tensor loads"
    assert count(r"\bLDTM\.") > 0, "no tcgen05.ld (tensor-memory epilogue)"
    assert count(r"\bUTCBAR") > 0, "no tcgen05.commit"
    assert count(r"\bHMMA\.") == 0 and count(r"\bHGMMA") == 0, "legacy mma.sync / wgmma code path present"


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not LIB.exists(), reason="needs cuobjdump and the built library")
def test_wide_backward_gemms_issue_two_mma_groups_per_k_step():
    """The 256 x 512 instantiations the backward GEMMs run on (gemm_sm100_kernel<2, 512, 4, ...>): eight cta_group::2
    tcgen05.mma per K block (4 K steps x 2 N = 256 groups sharing the A stage) against four in the 256 x 256 form, the
    d_weight epilogue through TMA reduce-add (UTMAREDG) and TMA store (UTMASTG), d_hidden and d_weight both present."""
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, timeout=300).stdout
    funcs = {}
    for chunk in sass.split("Function : ")[1:]:
        name, _, body = chunk.partition("\n")
        funcs[name.strip()] = body
    mma = lambda body: len(re.findall(r"\bUTCHMMA\.2CTA\b", body))
    wide = {n: b for n, b in funcs.items() if "gemm_sm100_kernelILi2ELi512ELi4E" in n}
    narrow = {n: b for n, b in funcs.items() if "gemm_sm100_kernelILi2ELi256ELi6E" in n}
    assert wide and narrow
    assert all(mma(b) == 8 for b in wide.values()), {n[-60:]: mma(b) for n, b in wide.items()}
    assert all(mma(b) == 4 for b in narrow.values()), {n[-60:]: mma(b) for n, b in narrow.items()}
    d_hidden = [b for n, b in wide.items() if "EpiStoreBF16" in n and "Lb0ELb1E" in n]   # A K-major, B MN-major
    d_weight = [b for n, b in wide.items() if "EpiTmaF32" in n and "Lb1ELb1E" in n]      # both MN-major
    assert len(d_hidden) == 1 and len(d_weight) == 1
    assert re.search(r"\bUTMAREDG\b", d_weight[0]) and re.search(r"\bUTMASTG\b", d_weight[0])
    assert not re.search(r"\bHMMA\.", d_hidden[0] + d_weight[0])
