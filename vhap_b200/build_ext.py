"""Builds vhap_b200/libvhap_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
SO = Path(os.environ.get("VH_SO_OUT", HERE / "libvhap_b200.so"))      # VH_SO_OUT / VH_EXTRA_FLAGS: experiment variants (dev only)
SOURCES = ["flame.cu", "raster.cu", "render.cu", "texture.cu", "blend_tc.cu", "dp_tex.cu", "api.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "--fmad=true"] + os.environ.get("VH_EXTRA_FLAGS", "").split()


def _newer(target: Path, deps) -> bool:
    return (not target.exists()) or any(d.stat().st_mtime > target.stat().st_mtime for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [HERE.parent / "include" / "vhap_b200.h"]
    if not force and not _newer(SO, deps):
        return SO
    objdir = HERE / ("build" if "VH_SO_OUT" not in os.environ else "build_" + SO.stem)
    objdir.mkdir(exist_ok=True)

    def cc(src):
        obj = objdir / (src + ".o")
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=6) as ex:
        objs = list(ex.map(cc, SOURCES))
    cmd = [NVCC, "-shared", "-o", str(SO)] + [str(o) for o in objs] + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
