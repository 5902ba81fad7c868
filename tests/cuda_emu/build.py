"""TEST INFRASTRUCTURE: builds tests/cuda_emu/build/libcuda_emu.so = a few of the library's .cu files compiled by g++ against the
CUDA-on-CPU shim in this directory (device code paths, one OS thread per CUDA thread).  Source transformations (text only,
into a scratch directory): kernel launches `k<<<cfg>>>(args);` -> emu::run(emu::cfg(cfg), [&]{ k(args); }); and dynamic
shared-memory declarations -> a static buffer."""
from __future__ import annotations

import re
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "orb_slam3_rgbl_b200" / "csrc"
BUILD = HERE / "build"
LIB = BUILD / "libcuda_emu.so"
KERNEL_FILES = ["orb_kernels.cu", "fast_strip_kernels.cu", "describe_warp_kernels.cu", "quadtree_kernels.cu", "pose_kernels.cu", "depth_kernels.cu", "depth_dilate_v2.cu"]
HEADERS = ["fast_strip.cuh", "describe_warp.cuh", "quadtree_block.cuh", "rgbl_device.cuh", "rgbl_kernels.h", "rgbl_internal.h",
           "orb_pattern_31.inc"]
HOST_FILES = ["host_tables.cpp", "quadtree_host.cpp"]


def _transform(text: str) -> str:
    text = re.sub(r"extern __shared__ __align__\(16\) (\w+(?: \w+)?) (\w+)\[\];", r"alignas(16) static \1 \2[emu::kDynSmem];", text)
    text = re.sub(r"(\w+)<<<(.+?)>>>\((.*?)\);", r"emu::run(emu::cfg(\2), [&]() { \1(\3); });", text, flags=re.S)
    # the two approximate FP64 MUFU seeds of the LM kernel (results have 32 zero low mantissa bits, PTX ISA "rcp.approx.ftz.f64")
    text = re.sub(r'asm\("rcp\.approx\.ftz\.f64 %0, %1;" : "=d"\((\w+)\) : "d"\((\w+)\)\);', r"\1 = emu::approx64(1.0 / \2);", text)
    text = re.sub(r'asm\("rsqrt\.approx\.ftz\.f64 %0, %1;" : "=d"\((\w+)\) : "d"\((\w+)\)\);', r"\1 = emu::approx64(1.0 / std::sqrt(\2));", text)
    text = text.replace('#include "../../include/rgbl_b200.h"', f'#include "{ROOT}/include/rgbl_b200.h"')
    return text


def build(force: bool = False) -> Path:
    srcs = [CSRC / f for f in KERNEL_FILES + HEADERS + HOST_FILES] + [HERE / "cuda_runtime.h", HERE / "emu_runtime.cpp", HERE / "emu_entry.cpp", Path(__file__)]
    if LIB.exists() and not force and all(LIB.stat().st_mtime > s.stat().st_mtime for s in srcs):
        return LIB
    if BUILD.exists():
        shutil.rmtree(BUILD)
    BUILD.mkdir(parents=True)
    for f in KERNEL_FILES + HEADERS + HOST_FILES:
        out = BUILD / (f[:-3] + ".emu.cpp" if f.endswith(".cu") else f)
        out.write_text(_transform((CSRC / f).read_text()))
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-pthread", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas", "-Wno-attributes", "-DPOSE_MIXED_SOLVE=1",
           f"-I{HERE}", f"-I{BUILD}", "-o", str(LIB), str(HERE / "emu_entry.cpp"), str(HERE / "emu_runtime.cpp"), *[str(BUILD / f) for f in HOST_FILES]]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
