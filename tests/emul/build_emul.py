"""TEST INFRASTRUCTURE: compile csrc/train_glue.cu, csrc/infer.cu and csrc/augment.cu for the HOST against tests/emul/cuda_emul.h
(see that header) into tests/emul/_build/librsb_emul.so.  The sources are used as they lie in csrc/ — two textual rewrites
only: the rsb_common.cuh include becomes the shim, and `kernel<<<grid, block, smem, stream>>>(args)` becomes
EMU_LAUNCH(cooperative?, kernel, grid, block, args)."""
from __future__ import annotations

import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "r-super_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "librsb_emul.so")
SOURCES = ["train_glue.cu", "infer.cu", "augment.cu", "seg_loss.cu", "morph.cu", "elementwise.cu", "stem_head.cu", "report_loss.cu", "medformer.cu"]
# kernels that use __syncthreads / warp shuffles: their blocks run as real threads
COOPERATIVE = {"grad_sqnorm_kernel", "clip_adamw_ema_kernel", "aug_stats_partial_kernel", "aug_stats_final_kernel",
               "seg_loss_pass1_kernel", "seg_loss_finalize_kernel", "seg_loss_pass2_kernel",
               # no barrier inside, but lock-free: run them as 256 truly concurrent threads so that the atomicMin union-find and
               # the flatten pass are exercised under real interleavings, not only under the sequential schedule
               "cc_merge_kernel", "cc_flatten_kernel"}

# files whose every kernel runs as real threads (shared-memory statistics flushes in helper functions)
ALL_COOPERATIVE = set()

LAUNCH = re.compile(r"(\w+(?:<[\w, ]+>)?)\s*<<<\s*([^;]*?)>>>\s*\(", re.S)


def _split_args(s: str):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "(<[":
            depth += 1
        elif ch in ")>]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


SYNC = re.compile(r"__syncthreads|__shfl_\w+_sync|__syncthreads_or|__ballot_sync|__any_sync")
FUNC = re.compile(r"\b(\w+)\s*\([^;{}()]*(?:\([^()]*\)[^;{}()]*)*\)\s*(?:const\s*)?\{")


def cooperative_functions(src: str) -> set:
    """Names of the functions (kernels and device helpers) of a translation unit that reach a barrier or a warp shuffle,
    directly or through a helper: those kernels must run as real threads.  Fixpoint over a rough brace-matched parse."""
    bodies = {}
    for m in FUNC.finditer(src):
        name = m.group(1)
        if name in ("if", "for", "while", "switch", "return", "sizeof"):
            continue
        i, depth = m.end(), 1
        while depth and i < len(src):
            depth += {"{": 1, "}": -1}.get(src[i], 0)
            i += 1
        bodies.setdefault(name, "")
        bodies[name] += src[m.end():i]
    coop = {n for n, b in bodies.items() if SYNC.search(b)} | {"warp_sum"}
    changed = True
    while changed:
        changed = False
        for n, b in bodies.items():
            if n not in coop and any(re.search(r"\b%s\s*(?:<[^;()]*>)?\s*\(" % re.escape(c), b) for c in coop):
                coop.add(n)
                changed = True
    return coop


def rewrite(src: str, all_coop: bool = False) -> str:
    auto = cooperative_functions(src)
    src = src.replace('#include "rsb_common.cuh"', '#include "rsb_common_emul.h"')
    # dynamic shared memory: `extern __shared__ float name[];` -> a pointer into the launch's buffer
    src = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];", r"\1* \2 = reinterpret_cast<\1*>(emu::g_dyn_smem.data());", src)
    pos, out = 0, ""
    for m in LAUNCH.finditer(src):
        kernel, cfg = m.group(1), _split_args(m.group(2))
        base = kernel.split("<")[0]
        # find the matching ')' of the argument list
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        args = src[m.end():i - 1]
        coop = base in COOPERATIVE or base in auto or all_coop
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += src[pos:m.start()] + f"EMU_LAUNCH({'true' if coop else 'false'}, ({kernel}), {cfg[0]}, {cfg[1]}, {smem}, {args})"
        pos = i
    return out + src[pos:]


def build() -> str:
    os.makedirs(OUT, exist_ok=True)
    h = hashlib.sha256()
    texts = []
    for s in SOURCES + [os.path.join(HERE, "cuda_emul.h"), os.path.join(HERE, "conv3_double.cpp"), os.path.abspath(__file__), "rsb_common.cuh"]:
        p = s if os.path.isabs(s) else os.path.join(CSRC, s)
        t = open(p).read()
        h.update(t.encode())
        texts.append(t)
    stamp = os.path.join(OUT, "stamp")
    if os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return LIB
    # the PTX-free tail of rsb_common.cuh (packing / vector access, warp_sum, statistics helpers) is used as it is
    common = open(os.path.join(CSRC, "rsb_common.cuh")).read()
    tail = common[common.index("// packing / vector access"):]
    with open(os.path.join(OUT, "rsb_common_emul.h"), "w") as f:
        f.write('#pragma once\n#include "cuda_emul.h"\nnamespace rsb {\n// ' + tail)
    cpps = []
    for name, text in zip(SOURCES, texts):
        cpp = os.path.join(OUT, name[:-3] + "_emul.cpp")
        with open(cpp, "w") as f:
            f.write(rewrite(text, all_coop=name in ALL_COOPERATIVE))
        cpps.append(cpp)
    with open(os.path.join(OUT, "api_emul.cpp"), "w") as f:
        f.write('#include "cuda_emul.h"\n#include "../../../include/rsuper_b200.h"\n'
                'extern "C" const char* rsb_last_error(void) { return rsb::g_last_error; }\n'
                'extern "C" int rsb_num_sms(void) { return 4; }\n')          # 4 "SMs": small grids keep the emulation quick
    cpps.append(os.path.join(OUT, "api_emul.cpp"))
    cpps.append(os.path.join(HERE, "conv3_double.cpp"))       # naive stand-in for the tensor-core kernels (see its header)
    cmd = ["g++", "-std=c++20", "-O2", "-g", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-attributes", "-I", HERE, "-I", OUT,
           "-o", LIB] + cpps
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"emulation build failed:\n{r.stdout}\n{r.stderr[-6000:]}")
    with open(stamp, "w") as f:
        f.write(h.hexdigest())
    return LIB


if __name__ == "__main__":
    print(build())
