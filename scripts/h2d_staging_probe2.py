"""Why is a pinned buffer filled by several host threads DMA'd slowly?  Variants of the fill, same 7.9 MB copy."""
import ctypes, glob, os, threading, time
import numpy as np
import torch

n = 7_878_656
torch.cuda.init(); torch.zeros(1, device="cuda")
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = ctypes.CDLL(cands[0])
def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags)) == 0
    return p.value, np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(p.value))
d = torch.empty(n, dtype=torch.uint8, device="cuda")
src = np.random.randint(0, 255, n, dtype=np.uint8)
def copy_ms(ptr):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    assert rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(n), ctypes.c_int(1), ctypes.c_void_p(0)) == 0
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
def fill_threads(arr, k):
    step = n // k
    th = [threading.Thread(target=lambda i=i: np.copyto(arr[i*step:(i+1)*step if i < k-1 else n], src[i*step:(i+1)*step if i < k-1 else n])) for i in range(k)]
    [t.start() for t in th]; [t.join() for t in th]
def run(label, ptr, arr, fill, post=lambda: None):
    ts, fs = [], []
    for _ in range(12):
        t0 = time.perf_counter(); fill(); fs.append(time.perf_counter() - t0)
        post()
        ts.append(copy_ms(ptr))
    print(f"{label:46s} fill {np.median(fs)*1e3:6.3f} ms   H2D {np.median(ts):6.3f} ms ({n/np.median(ts)/1e6:5.1f} GB/s)", flush=True)
scratch = np.empty(n, dtype=np.uint8)
for name, flags in (("default", 0), ("write-combined", 4)):
    ptr, arr = host_alloc(n, flags)
    run(f"{name}: 1 thread", ptr, arr, lambda: np.copyto(arr, src))
    for k in (2, 4, 8):
        run(f"{name}: {k} threads", ptr, arr, lambda k=k: fill_threads(arr, k))
    run(f"{name}: 8 threads + 20 ms settle", ptr, arr, lambda: fill_threads(arr, 8), lambda: time.sleep(0.02))
    if flags == 0:
        run(f"{name}: 8 threads + main thread re-reads", ptr, arr, lambda: fill_threads(arr, 8), lambda: arr.sum())
        run(f"{name}: 8 threads fill scratch, main copies", ptr, arr, lambda: (fill_threads(scratch, 8), np.copyto(arr, scratch)) if True else None)
