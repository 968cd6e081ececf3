"""Build a host-emulation library of a functor-style engine: the SAME source as the CUDA build (nabladft_b200/csrc/<name>.cu) compiled as
plain C++ with -DNB_EMU, every kernel functor run as an (OpenMP) loop.  TEST INFRASTRUCTURE ONLY -- see emu_shim.h."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build(force: bool = False, name: str = "gemnet_oc") -> str:
    """name: a functor-style source of nabladft_b200/csrc (gemnet_oc, schnet_train)."""
    SRC = os.path.join(ROOT, "nabladft_b200", "csrc", name + ".cu")
    OUT = os.path.join(HERE, "_build", f"lib{name}_emu.so")
    import glob

    deps = glob.glob(os.path.join(ROOT, "nabladft_b200", "csrc", "*.cuh")) + glob.glob(os.path.join(ROOT, "nabladft_b200", "csrc", "*.inc")) + [SRC, os.path.join(ROOT, "nabladft_b200", "csrc", "gemnet_pf.cuh"), os.path.join(HERE, "emu_shim.h"), os.path.join(ROOT, "include", "nabla_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) < os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    tmp = f"{OUT}.{os.getpid()}.tmp"  # pytest-xdist workers may build at the same time: never expose a half-written library
    subprocess.check_call(["g++", "-O3", "-fopenmp", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-DNB_EMU", "-I", HERE, "-Wno-unknown-pragmas", SRC, "-o", tmp])
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    import sys

    print(build(force=True, name=sys.argv[1] if len(sys.argv) > 1 else "gemnet_oc"))
