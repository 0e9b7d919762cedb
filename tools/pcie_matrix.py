"""Host<->device bandwidth matrix of the box (VERDICT r1 task 4): pinned cudaMemcpyAsync H2D-only /
D2H-only / both ways, on 1, 2, 4, 8 GPUs at once, from ONE process (a thread per GPU) and from N
processes, with default and write-combined|portable pinned memory.  Pure ctypes on libcudart — none
of this repo's code is involved, so the table says what the machine gives.

usage: python tools/pcie_matrix.py [--gpus 1,2,4,8] [--mb 256] [--iters 8]       (driver)
       python tools/pcie_matrix.py --worker DEV ...                               (internal)
"""
import argparse
import ctypes as C
import glob
import json
import os
import subprocess
import sys
import threading
import time


def cudart_path():
    import importlib.util
    spec = importlib.util.find_spec("torch")  # only to locate the bundled libcudart, torch is not imported
    libdir = os.path.join(os.path.dirname(spec.origin), "lib")
    site = os.path.dirname(os.path.dirname(spec.origin))
    cands = (glob.glob(os.path.join(libdir, "libcudart*.so*")) +
             glob.glob(os.path.join(site, "nvidia", "cuda_runtime", "lib", "libcudart.so*")) +
             glob.glob("/usr/local/cuda/lib64/libcudart.so*"))
    return cands[0]


def cudart(path=None):
    lib = C.CDLL(path or cudart_path())
    lib.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    lib.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    lib.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    lib.cudaStreamCreateWithFlags.argtypes = [C.POINTER(C.c_void_p), C.c_uint]
    lib.cudaStreamSynchronize.argtypes = [C.c_void_p]
    return lib


class Dev:
    """Buffers and streams of one GPU."""

    def __init__(self, lib, dev, nbytes, flags):
        self.lib, self.dev, self.n = lib, dev, nbytes
        assert lib.cudaSetDevice(dev) == 0
        self.h = [C.c_void_p(), C.c_void_p()]
        self.d = [C.c_void_p(), C.c_void_p()]
        self.s = [C.c_void_p(), C.c_void_p()]
        for i in range(2):
            assert lib.cudaHostAlloc(C.byref(self.h[i]), nbytes, flags) == 0
            C.memset(self.h[i], 1, nbytes)
            assert lib.cudaMalloc(C.byref(self.d[i]), nbytes) == 0
            assert lib.cudaStreamCreateWithFlags(C.byref(self.s[i]), 1) == 0

    def run(self, kind, iters):
        lib = self.lib
        lib.cudaSetDevice(self.dev)
        for _ in range(iters):
            if kind in ("h2d", "both"):
                lib.cudaMemcpyAsync(self.d[0], self.h[0], self.n, 1, self.s[0])
            if kind in ("d2h", "both"):
                lib.cudaMemcpyAsync(self.h[1], self.d[1], self.n, 2, self.s[1])
        lib.cudaStreamSynchronize(self.s[0])
        lib.cudaStreamSynchronize(self.s[1])


def worker(args):
    """One process, one GPU: waits for the go-file, runs, prints seconds."""
    lib = cudart(args.lib or None)
    d = Dev(lib, args.worker, args.mb << 20, args.flags)
    d.run(args.kind, 1)
    print("ready", flush=True)
    while not os.path.exists(args.go):
        time.sleep(0.0005)
    t0 = time.perf_counter()
    d.run(args.kind, args.iters)
    print(json.dumps({"dev": args.worker, "s": time.perf_counter() - t0}), flush=True)


def one_process(lib, devs, kind, iters):
    bar = threading.Barrier(len(devs) + 1)
    out = {}

    def th(d):
        d.run(kind, 1)
        bar.wait()
        t0 = time.perf_counter()
        d.run(kind, iters)
        out[d.dev] = time.perf_counter() - t0
    ts = [threading.Thread(target=th, args=(d,)) for d in devs]
    for t in ts:
        t.start()
    bar.wait()
    for t in ts:
        t.join()
    return out


def n_processes(n, kind, args, flags):
    go = f"/tmp/pcie_go_{os.getpid()}_{kind}_{n}_{flags}"
    ps = [subprocess.Popen([sys.executable, __file__, "--worker", str(d), "--kind", kind, "--mb", str(args.mb),
                            "--iters", str(args.iters), "--flags", str(flags), "--go", go, "--lib", cudart_path()],
                           stdout=subprocess.PIPE, text=True) for d in range(n)]
    for p in ps:
        assert p.stdout.readline().strip() == "ready"
    open(go, "w").close()
    out = {}
    for p in ps:
        r = json.loads(p.stdout.readline())
        out[r["dev"]] = r["s"]
        p.wait()
    os.unlink(go)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--worker", type=int, default=-1)
    ap.add_argument("--kind", default="both")
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--go", default="")
    ap.add_argument("--lib", default="")
    ap.add_argument("--procs", default="2,8", help="GPU counts for the N-process runs")
    args = ap.parse_args()
    if args.worker >= 0:
        return worker(args)
    lib = cudart()
    cnt = C.c_int()
    lib.cudaGetDeviceCount(C.byref(cnt))
    counts = [int(x) for x in args.gpus.split(",") if int(x) <= cnt.value]
    gb = (args.mb << 20) * args.iters / 1e9
    print(f"{cnt.value} GPUs visible; {args.mb} MiB per copy x {args.iters}; GB/s per GPU (min..max) and aggregate, "
          "per direction")
    # cudaHostAllocDefault = 0; Portable = 1; WriteCombined = 4
    for flags, fname in ((0, "pinned default"), (1 | 4, "pinned write-combined|portable")):
        devs = [Dev(lib, d, args.mb << 20, flags) for d in range(max(counts))]
        for n in counts:
            for kind in ("h2d", "d2h", "both"):
                t = one_process(lib, devs[:n], kind, args.iters)
                rates = [gb / s for s in t.values()]
                print(f"{fname:32s} 1 process, {n} GPU(s) {kind:5s}: {min(rates):6.1f}..{max(rates):6.1f} per GPU, "
                      f"aggregate {sum(rates):7.1f}")
        del devs
        for n in [int(x) for x in args.procs.split(",") if x and int(x) <= cnt.value and int(x) > 1]:
            for kind in ("h2d", "d2h", "both"):
                t = n_processes(n, kind, args, flags)
                rates = [gb / s for s in t.values()]
                print(f"{fname:32s} {n} processes, {n} GPU(s) {kind:5s}: {min(rates):6.1f}..{max(rates):6.1f} per GPU, "
                      f"aggregate {sum(rates):7.1f}")


if __name__ == "__main__":
    main()
