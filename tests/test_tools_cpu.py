"""CPU checks of the developer tools that need no GPU: the static SASS timing model (tools/sass_chain.py) decodes the control
fields of the built LunarLander step kernel and finds its solver loops."""
import shutil
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not available")
def test_sass_chain_decodes_the_step_kernel(native_lib):
    import sass_chain as sc
    obj = ROOT / "gymrl_b200" / "lib" / "obj" / "env_lunar.o"
    assert obj.exists()
    text = sc.disasm(str(obj))
    for variant, lo, hi in ((0, 12000, 14000), (3, 18000, 21000)):
        code = sc.parse(text, f"lunar_step_kernelILi{variant}E")
        assert lo < len(code) < hi, len(code)
        # every instruction carries a decodable control word: stall count 0..15, barrier indices 0..7, 6-bit wait mask
        assert all(0 <= c["stall"] <= 15 and 0 <= c["wr"] <= 7 and 0 <= c["rd"] <= 7 and 0 <= c["wait"] < 64 for c in code)
        # loads set a write barrier (variable latency), plain FP32 arithmetic does not
        ldl = [c for c in code if c["op"].startswith("LDL")]
        ffma = [c for c in code if c["op"] == "FFMA"]
        assert ldl and all(c["wr"] != 7 for c in ldl)
        assert ffma and all(c["wr"] == 7 for c in ffma)
        loops = sc.loops(code)
        # the 180-iteration velocity loop(s) and the 60-iteration position loop are among the large loops
        big = [l for l in loops if l[2] > 1000]
        assert len(big) >= 3
        # replaying a loop body costs at least one cycle per instruction and at most the stall sum plus the modelled waits
        t, e, n, stall_sum, _ = max(loops, key=lambda l: l[2])
        cyc, issued, waited = sc.replay(code, t, e)
        assert issued >= 1 and cyc >= issued
    # arrangement 3 carries div_chain's inline reciprocal sequences next to the plain divisions of its repeat path
    r0 = sum(c["op"].startswith("MUFU.RCP") for c in sc.parse(text, "lunar_step_kernelILi0E"))
    r3 = sum(c["op"].startswith("MUFU.RCP") for c in sc.parse(text, "lunar_step_kernelILi3E"))
    assert r3 > r0 > 0
