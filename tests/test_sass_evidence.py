"""CPU: static evidence from the built library (cuobjdump; nvcc cross-compiles without a GPU) that the wide-state
kernels are what DESIGN.md says they are: tensor-core contractions (HMMA.1688.F32.TF32 = mma.sync.m16n8k8 tf32),
within the 128-register budget of a 512-thread CTA without meaningful spills, and with the operand split that
keeps the contraction loop at a handful of instructions per mma (the first version spent 7 instructions per
split on the cvt.rna.tf32 emulation: profiles/README.md)."""
import collections
import re
import shutil
import subprocess

import pytest

KERNELS = {"fwd": "_ZN3eng15fwd_wide_kernelENS_7FwdArgsE", "bwd": "_ZN3eng15bwd_wide_kernelENS_7BwdArgsE"}

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="needs the CUDA toolkit")


def _lib_path():
    import hgnn_b200  # noqa: F401
    from hgnn_b200 import _lib
    return _lib.LIB_PATH


def _sass(fn):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, _lib_path()], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())))
    return ins


def test_wide_kernels_fit_the_register_file():
    out = subprocess.run(["cuobjdump", "-res-usage", _lib_path()], capture_output=True, text=True).stdout
    for fn in KERNELS.values():
        m = re.search(re.escape(fn) + r":\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", out)
        assert m, "kernel %s not in the library" % fn
        reg, stack, shared = map(int, m.groups())
        assert reg <= 128            # 512 threads per CTA: 65536 / 512
        assert stack <= 16           # no spills to speak of
        assert shared <= 14 * 1024   # static part; the dispatch leaves 15 KB under the 227 KB limit for it


@pytest.mark.parametrize("which", ["fwd", "bwd"])
def test_wide_contractions_run_on_tensor_cores(which):
    ins = _sass(KERNELS[which])
    assert ins, "no SASS for %s" % KERNELS[which]
    hmma = [t for _, t in ins if t.startswith("HMMA")]
    assert len(hmma) >= 18 and all(t.startswith("HMMA.1688.F32.TF32") for t in hmma)
    # innermost loops that contain mma: instructions per mma
    best = None
    for addr, t in ins:
        m = re.match(r"BRA\s+(0x[0-9a-f]+)", t)
        if not m or int(m.group(1), 16) >= addr:
            continue
        body = [x for a, x in ins if int(m.group(1), 16) <= a <= addr]
        n = sum(x.startswith("HMMA") for x in body)
        if n >= 6 and (best is None or len(body) < len(best)):
            best = body
    assert best is not None, "no loop with mma found"
    n = sum(x.startswith("HMMA") for x in best)
    ops = collections.Counter(x.split()[0].split(".")[0] for x in best)
    assert len(best) / n <= 8.0, (len(best), n, ops.most_common(8))
    # the cvt.rna.tf32 emulation (FSETP against +INF, SEL) must not be back in the loop
    assert ops["FSETP"] == 0 and ops["SEL"] == 0, ops.most_common(8)


def test_tcgen05_forward_kernel_is_blackwell_native():
    """csrc/engine_tc5.cuh (opt-in, HGNN_B200_WIDE_TC5=1; parity-tested on the GPU): the contraction is issued as
    tcgen05.mma (UTCHMMA) with the accumulator in tensor memory (LDTM = tcgen05.ld) and completion through an
    mbarrier (UTCBAR = tcgen05.commit) - not as warp-level mma.sync."""
    ins = _sass("_ZN3eng14fwd_tc5_kernelENS_7FwdArgsE")
    assert ins, "fwd_tc5_kernel not in the library"
    ops = collections.Counter(t.split()[0].split(".")[0] for _, t in ins)
    assert ops["UTCHMMA"] >= 3 and ops["LDTM"] >= 1 and ops["UTCBAR"] >= 1, ops.most_common(12)
    assert ops["HMMA"] == 0


H2_KERNELS = {
    # the default width-4 (h = 2) training kernels of the C2 / C4 steps: name -> (max registers, max stack bytes)
    "_ZN3eng16bwd_row4p_kernelILi2ELi4ELb0EEEvNS_8Bwd4ArgsE": (128, 16),      # 4 CTAs of 128 threads per SM
    "_ZN3eng16bwd_row4p_kernelILi2ELi2ELb0EEEvNS_8Bwd4ArgsE": (128, 16),
    "_ZN3eng15fwd_row4_kernelILi1ELb1ELi8ELi4EEEvNS_8Fwd4ArgsE": (170, 16),    # 3 CTAs per SM
    "_ZN3eng15fwd_row4_kernelILi1ELb1ELi8ELi16EEEvNS_8Fwd4ArgsE": (255, 32),   # 2 CTAs per SM
}


def test_default_h2_kernels_keep_their_register_budget():
    """The thread-per-row kernels are latency bound: their CTAs-per-SM (registers) and the absence of spills in the row
    loop are what the measured step time rests on; a change that silently adds 10 registers to the
    pipelined backward drops it from 4 to 3 CTAs per SM (measured: 0.59 -> 0.68 ms per step)."""
    out = subprocess.run(["cuobjdump", "-res-usage", _lib_path()], capture_output=True, text=True).stdout
    for fn, (max_reg, max_stack) in H2_KERNELS.items():
        m = re.search(re.escape(fn) + r":\s*\n\s*REG:(\d+) STACK:(\d+)", out)
        assert m, "kernel %s not in the library" % fn
        reg, stack = map(int, m.groups())
        assert reg <= max_reg and stack <= max_stack, (fn, reg, stack)


def test_default_h2_kernels_use_dependent_launch_and_vector_reductions():
    """Programmatic dependent launch (griddepcontrol.wait = ACQBULK) in the forward and the pipelined backward; the
    backward accumulates into gX with red.global.add.v4.f32 (REDG.E.ADD.F32x4) and flushes its sums with fp64 REDs."""
    for fn in H2_KERNELS:
        ops = collections.Counter(t.split()[0] for _, t in _sass(fn))
        assert ops["ACQBULK"] >= 1, (fn, "no griddepcontrol.wait")
        assert any(k.startswith("REDG.E.ADD.F64") for k in ops), (fn, "no fp64 reductions")
        if "bwd_row4p" in fn:
            assert any(k.startswith("REDG.E.ADD.F32x4") for k in ops), (fn, "no vector reduction into gX")
