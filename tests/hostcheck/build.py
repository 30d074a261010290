"""Builds tests/hostcheck/libhostcheck.so (g++, host only).  Developer aid for checking the device math headers
against the oracle without a GPU; not part of the product library."""
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent


def build(force=False):
    so = HERE / "libhostcheck.so"
    srcs = [HERE / "hostcheck.cpp"] + sorted((HERE.parents[1] / "vhap_b200" / "csrc").glob("*.cuh"))
    if so.exists() and not force and all(so.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return so
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DVH_HOST_CHECK",
                           str(HERE / "hostcheck.cpp"), "-o", str(so)])
    return so


if __name__ == "__main__":
    print(build(True))
