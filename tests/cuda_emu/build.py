"""TEST INFRASTRUCTURE: compile a simple CUDA source of rmem_b200/csrc for the host emulation of tests/cuda_emu/cuda_emu.h.

    lib = build("train_loss.cu")     # -> ctypes.CDLL of the emulated translation unit (its extern "C" entry points)

The source is used as it is; only the launch statements `kernel<<<grid, block, 0, stream>>>(args)` are rewritten to
emu::launch(kernel, grid, block, args).  Compiled with -ffp-contract=off so that the fp32 arithmetic is the unfused
round-to-nearest sequence the kernels spell out with __fmul_rn / __fadd_rn.  Pointers passed to the entry points are host
pointers.  Only tests/ use this; nothing of it ships."""
import ctypes
import hashlib
import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "rmem_b200", "csrc")
LAUNCH = re.compile(r"(\w+)<<<\s*([^,<>]+?)\s*,\s*([^,<>]+?)\s*,\s*0\s*,\s*([^,<>]+?)\s*>>>\(")

SUPPORT = """
#include <cstdarg>
namespace rmem {
static thread_local char g_err[1024] = "";
static thread_local long long g_launches = 0;
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); }
const char* get_error() { return g_err; }
long long& launch_counter() { return g_launches; }
}
extern "C" const char* rmem_last_error(void) { return rmem::get_error(); }
"""


DYN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?unsigned char (\w+)\[\];")


def build(source: str, extra: str = "") -> ctypes.CDLL:
    """`extra`: C++ appended to the translation unit (extern "C" shims over its host functions for the tests)."""
    src = open(os.path.join(CSRC, source)).read()
    out, n = LAUNCH.subn(r"emu::launch(\1, dim3(\2), dim3(\3), ", src)
    assert n == src.count("<<<"), f"{source}: {n} of {src.count('<<<')} launches rewritten"
    out = DYN_SMEM.sub(r"unsigned char* \1 = emu::ctx.dyn_smem;", out)
    out = out.replace('#include "ops.cuh"', f'#include "{CSRC}/ops.cuh"')
    out = out.replace('#include "../../include/rmem_b200.h"', f'#include "{ROOT}/include/rmem_b200.h"')
    out = out.replace('#include "common.cuh"', f'#include "{CSRC}/common.cuh"')
    out += extra
    hdrs = "".join(open(os.path.join(d, f)).read() for d in (HERE, os.path.join(HERE, "stubs")) for f in sorted(os.listdir(d))
                   if f.endswith(".h"))
    tag = hashlib.sha256((out + SUPPORT + hdrs).encode()).hexdigest()[:16]
    cache = os.path.join(tempfile.gettempdir(), "rmem_cuda_emu")
    os.makedirs(cache, exist_ok=True)
    so = os.path.join(cache, f"{source}.{tag}.so")
    if not os.path.exists(so):
        cpp = os.path.join(cache, f"{source}.{tag}.cpp")
        with open(cpp, "w") as fh:
            fh.write(out + SUPPORT)
        cmd = ["g++", "-std=c++20", "-O1", "-g", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-x", "c++",
               "-I", os.path.join(HERE, "stubs"), "-include", os.path.join(HERE, "cuda_emu.h"), cpp, "-o", so + ".tmp"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"host emulation build of {source} failed:\n{r.stderr[-4000:]}")
        os.replace(so + ".tmp", so)
    lib = ctypes.CDLL(so)
    lib.rmem_last_error.restype = ctypes.c_char_p
    return lib
