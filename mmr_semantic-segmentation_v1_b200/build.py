"""Build libmmrseg.so (sm_100a only) in-tree with nvcc.

`python -m mmrseg_b200.build` or `__graft_entry__.build()`.  The library links cudart
statically and resolves the driver's cuTensorMapEncodeTiled at run time, so it loads
(and exports its C-ABI) on a box with no GPU; nvcc cross-compiles there.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmmrseg.so")
SOURCES = ["common.cu", "conv_gemm.cu", "conv_halo.cu", "conv_wgrad.cu", "conv_wgrad_halo.cu", "elementwise.cu", "loss_metric.cu",
           "optim.cu", "sliding_window.cu", "hausdorff.cu", "pointwise_head.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-DMMR_BUILD",
]


def _stamp():
    h = hashlib.sha256()
    for root, _, files in sorted(os.walk(CSRC)):
        for f in sorted(files):
            with open(os.path.join(root, f), "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    with open(os.path.join(HERE, "..", "include", "mmrseg.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp_file = LIB + ".stamp"
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file):
        if open(stamp_file).read().strip() == stamp:
            return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        text = out.decode(errors="replace")
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, text))
        elif verbose or "warning" in text:
            sys.stderr.write(text)
    if failed:
        raise RuntimeError("libmmrseg.so build failed")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-cudart", "static"]
    subprocess.check_call(cmd)
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return LIB


def build_variant(name, extra_flags):
    """A/B builds: libmmrseg_<name>.so with extra nvcc flags (e.g. -DMMR_PDL_LATE=0), loaded through MMR_LIB."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    out_dir = os.path.join(HERE, "build", name)
    os.makedirs(out_dir, exist_ok=True)
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(out_dir, src.replace(".cu", ".o"))
        procs.append(subprocess.Popen([nvcc] + NVCC_FLAGS + list(extra_flags) + ["-c", os.path.join(CSRC, src), "-o", obj]))
        objs.append(obj)
    if any(p.wait() != 0 for p in procs):
        raise RuntimeError("variant build failed")
    lib = os.path.join(HERE, "libmmrseg_%s.so" % name)
    subprocess.check_call([nvcc, "-shared", "-o", lib] + objs + ["-cudart", "static"])
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
