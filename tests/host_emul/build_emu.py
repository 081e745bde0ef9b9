"""tests/host_emul/build_emu.py — TEST INFRASTRUCTURE.

Builds `tests/host_emul/_build/libmgb200_emu.so`: the product's own sources
(multigrid_nikhil_c-_b200/csrc/*.cu, *.cuh, *.h — host orchestration AND kernel bodies) compiled with g++ against
the CUDA-on-CPU emulation in tests/host_emul/cuda_emu/.  The CPU test-suite loads it in place of libmgb200.so
(tests/conftest.py, MGB200_TEST_EMU=1) to run the GPU parity tests without a GPU.  The product never loads it.

Source rewriting (everything else is handled by the emulation headers):
  kernel<T...><<<grid, block, smem, stream>>>(args);   ->  ::emu::launch(kernel<T...>, dim3(grid), dim3(block), smem, stream, args);
  extern __shared__ __align__(16) unsigned char name[]; ->  unsigned char* name = ::emu::dyn_smem();
  __noinline__                                          ->  EMU_NOINLINE   (libstdc++ uses the bare token in attributes)
"""
from __future__ import annotations

import hashlib
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "multigrid_nikhil_c-_b200", "csrc")
EMU = os.path.join(HERE, "cuda_emu")
BUILD = os.path.join(HERE, "_build")
OUT = os.path.join(BUILD, "libmgb200_emu.so")
UNITS = ["capi.cu", "ctx.cu", "comm.cu", "fused.cu"]

_LAUNCH = re.compile(r"(\bk_\w+(?:<[^<>;]*>)?)\s*<<<(.*?)>>>\s*\((.*?)\);")
_SHARED = re.compile(r"extern\s+__shared__\s+__align__\(\d+\)\s+unsigned\s+char\s+(\w+)\[\];")


def _split_top(s: str) -> list:
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite(text: str, name: str) -> str:
    def launch(m):
        kern, cfg, args = m.group(1), _split_top(m.group(2)), m.group(3).strip()
        if len(cfg) != 4:
            raise RuntimeError(f"{name}: launch configuration with {len(cfg)} parameters: {m.group(0)}")
        grid, block, smem, stream = cfg
        return f"::emu::launch({kern}, dim3({grid}), dim3({block}), {smem}, {stream}{', ' + args if args else ''});"

    text, n = _LAUNCH.subn(launch, text)
    if "<<<" in text:
        raise RuntimeError(f"{name}: a kernel launch was not rewritten")
    text = _SHARED.sub(r"unsigned char* \1 = ::emu::dyn_smem();", text)
    if re.search(r"extern\s+__shared__", text):
        raise RuntimeError(f"{name}: an extern __shared__ declaration was not rewritten")
    return text.replace("__noinline__", "EMU_NOINLINE")


def _sources() -> list:
    files = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    return [os.path.join(CSRC, f) for f in files]


def _digest() -> str:
    h = hashlib.sha256()
    paths = _sources() + [os.path.join(EMU, f) for f in sorted(os.listdir(EMU))] + [os.path.abspath(__file__),
                                                                                   os.path.join(ROOT, "include", "mgb200.h")]
    for p in paths:
        h.update(p.encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = os.path.join(BUILD, "stamp")
    dig = _digest() + os.environ.get("CUDA_EMU_UBSAN", "")
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    src_dir = os.path.join(BUILD, "pkg", "csrc")     # same depth as the real csrc: "../../include/mgb200.h" resolves
    shutil.rmtree(os.path.join(BUILD, "pkg"), ignore_errors=True)
    os.makedirs(src_dir, exist_ok=True)
    os.makedirs(os.path.join(BUILD, "include"), exist_ok=True)
    shutil.copy(os.path.join(ROOT, "include", "mgb200.h"), os.path.join(BUILD, "include", "mgb200.h"))
    for p in _sources():
        base = os.path.basename(p)
        out_name = base[:-3] + ".cpp" if base.endswith(".cu") else base
        with open(os.path.join(src_dir, out_name), "w") as f:
            f.write(rewrite(open(p).read(), base))
    flags = ["-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-fno-strict-aliasing", "-w", "-DMGB_EMU=1", "-I" + EMU]
    # CUDA_EMU_UBSAN=1: alignment / bounds / integer-overflow checks in every kernel and host function (a misaligned
    # 16-byte vector access is a fault on the GPU but usually silent on x86)
    san = os.environ.get("CUDA_EMU_UBSAN") == "1"
    if san:
        flags += ["-fsanitize=alignment,bounds,signed-integer-overflow,shift,null", "-fno-sanitize-recover=all", "-g"]
    objs, procs = [], []
    for u in UNITS + ["emu_runtime.cpp"]:
        src = os.path.join(EMU, u) if u == "emu_runtime.cpp" else os.path.join(src_dir, u[:-3] + ".cpp")
        obj = os.path.join(BUILD, os.path.splitext(u)[0] + ".o")
        objs.append(obj)
        procs.append((u, subprocess.Popen(["g++", *flags, "-c", src, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for u, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"emulation build of {u} failed:\n{out[-6000:]}")
        if verbose and out:
            print(out)
    link = subprocess.run(["g++", "-shared", "-o", OUT, *objs, "-ldl", *(["-fsanitize=undefined"] if san else [])], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if link.returncode != 0:
        raise RuntimeError("emulation link failed:\n" + link.stdout[-4000:])
    with open(stamp, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
