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
