"""How fast can pageable memory be pinned IN PLACE (cudaHostRegister) and DMA'd from there, compared with the staging ring
(memcpy into pinned lanes)?  One GPU.  Decides whether an in-place path is worth building for multi-GPU ingest, where
the ring's extra pass over host DRAM is what saturates (8 ranks: 67 GB/s aggregate on the round-2 box)."""
import ctypes as C, os, sys, time, threading, numpy as np
sys.path.insert(0, ".")
import smartcore_b200 as sc
rt = C.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else C.CDLL("libcudart.so")
rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]; rt.cudaHostUnregister.argtypes = [C.c_void_p]
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]; rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaStreamCreate.argtypes = [C.POINTER(C.c_void_p)]; rt.cudaStreamSynchronize.argtypes = [C.c_void_p]; rt.cudaSetDevice.argtypes = [C.c_int]
ctx = sc.Context(0)
n = 5_120_000_000
x = np.empty(n, dtype=np.uint8); x[:] = 1
base = x.ctypes.data
dev = C.c_void_p(); assert rt.cudaMalloc(C.byref(dev), n) == 0
print("cores", os.cpu_count())
for chunk_mb in (64, 256):
    for T in (1, 2, 4, 8):
        chunk = chunk_mb << 20
        nchunks = (n + chunk - 1) // chunk
        t_reg = [0.0] * T
        def work(t):
            rt.cudaSetDevice(0)
            st = C.c_void_p(); rt.cudaStreamCreate(C.byref(st))
            for c in range(t, nchunks, T):
                off = c * chunk; ln = min(chunk, n - off)
                a = time.perf_counter()
                rc = rt.cudaHostRegister(base + off, ln, 0)
                t_reg[t] += time.perf_counter() - a
                if rc: print("register failed", rc); return
                rt.cudaMemcpyAsync(dev.value + off, base + off, ln, 1, st)
                rt.cudaStreamSynchronize(st)
                rt.cudaHostUnregister(base + off)
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
        [t.start() for t in th]; [t.join() for t in th]
        dt = time.perf_counter() - t0
        print("register+DMA+unregister in place: chunk %4d MB, %d threads: %.3f s = %5.1f GB/s (register alone %.3f s per thread)" % (
            chunk_mb, T, dt, n / dt / 1e9, max(t_reg)), flush=True)
xs = x.view(np.float64).reshape(-1, 64)
for _ in range(2):
    t0 = time.perf_counter(); ds = ctx.upload(xs); dt = time.perf_counter() - t0; ds.close()
    print("staging ring (library): %.3f s = %.1f GB/s" % (dt, n / dt / 1e9), flush=True)
