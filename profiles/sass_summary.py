"""profiles/sass_summary.py — static evidence from the built library (no GPU needed):
per kernel, registers / spills / shared memory from `ptxas -v` (multigrid_nikhil_c-_b200/lib/obj/*.ptxas.log) and
instruction-mix counts from `cuobjdump -sass` (128-bit global accesses, LDGSTS = cp.async, warp shuffles, FP64 adds /
multiplies vs fused multiply-adds -- the library is built with --fmad=false, the parity contract -- and the absence of
tensor-core instructions: the path is bandwidth-bound stencil work).
    python profiles/sass_summary.py > profiles/r01_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multigrid_nikhil_c-_b200", "lib", "libmgb200.so")
OBJ = os.path.join(ROOT, "multigrid_nikhil_c-_b200", "lib", "obj")


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"mgb::", "", re.sub(r"\(.*\)$", "", re.sub(r"^void ", "", n))) for n in out]


def ptxas():
    info = {}
    for f in sorted(os.listdir(OBJ)):
        if not f.endswith(".ptxas.log"):
            continue
        txt = open(os.path.join(OBJ, f)).read()
        for blk in re.split(r"ptxas info\s+: Compiling entry function '", txt)[1:]:
            name = blk.split("'")[0]
            regs = re.search(r"Used (\d+) registers", blk)
            spill = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", blk)
            smem = re.search(r"(\d+) bytes smem", blk)
            info[name] = (int(regs.group(1)) if regs else -1, int(spill.group(2)) if spill else 0, int(smem.group(1)) if smem else 0)
    return info


def sass():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    mix = {}
    for blk in re.split(r"\n\s*Function : ", txt)[1:]:
        name = blk.split("\n")[0].strip()
        ops = collections.Counter()
        for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", blk, flags=re.M):
            ops[m.group(1)] += 1
        mix[name] = ops
    return mix


def count(ops, pred):
    return sum(v for k, v in ops.items() if pred(k))


def main():
    reg, mix = ptxas(), sass()
    names = sorted(mix)
    pretty = dict(zip(names, demangle(names)))
    print("# Static SASS / ptxas summary of libmgb200.so (sm_100a, nvcc 12.9, -O3 --fmad=false); produced by profiles/sass_summary.py")
    print("# cols: regs spill_B static_smem_B | instr LDG.128 STG.128 LDGSTS(cp.async) UBLKCP(TMA) SYNCS(mbarrier) LDS STS SHFL DADD DMUL DFMA FADD FMUL FFMA BAR tensor(HMMA/UTC*MMA)")
    tot_tensor = 0
    for n in sorted(names, key=lambda k: pretty[k]):
        o = mix[n]
        r = reg.get(n, (-1, 0, 0))
        tensor = count(o, lambda k: "MMA" in k)
        tot_tensor += tensor
        print(f"{pretty[n][:78]:78s} {r[0]:4d} {r[1]:4d} {r[2]:6d} | {sum(o.values()):6d} "
              f"{count(o, lambda k: k.startswith('LDG') and '.128' in k and not k.startswith('LDGSTS')):4d} "
              f"{count(o, lambda k: k.startswith('STG') and '.128' in k):4d} "
              f"{count(o, lambda k: k.startswith('LDGSTS')):4d} "
              f"{count(o, lambda k: k.startswith('UBLKCP')):4d} {count(o, lambda k: k.startswith('SYNCS')):4d} "
              f"{count(o, lambda k: k.startswith('LDS')):4d} {count(o, lambda k: k.startswith('STS')):4d} "
              f"{count(o, lambda k: k.startswith('SHFL')):4d} "
              f"{count(o, lambda k: k.startswith('DADD')):4d} {count(o, lambda k: k.startswith('DMUL')):4d} {count(o, lambda k: k.startswith('DFMA')):4d} "
              f"{count(o, lambda k: k.startswith('FADD')):4d} {count(o, lambda k: k.startswith('FMUL')):4d} {count(o, lambda k: k.startswith('FFMA')):4d} "
              f"{count(o, lambda k: k.startswith('BAR') or k.startswith('UCGABAR')):4d} {tensor:3d}")
    print(f"# kernels: {len(names)}; tensor-core instructions in the whole library: {tot_tensor}")


if __name__ == "__main__":
    main()
