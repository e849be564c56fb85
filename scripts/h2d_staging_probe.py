import torch, time, threading, numpy as np
n = 7_878_656
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
src = np.random.randint(0, 255, n, dtype=np.uint8)
hn = h.numpy()
def timed(label, fill):
    ts = []
    for _ in range(10):
        fill()
        time.sleep(0.002)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{label}: H2D {np.median(ts):.3f} ms ({n/np.median(ts)/1e6:.1f} GB/s)")
timed("no refill", lambda: None)
timed("refill main thread", lambda: np.copyto(hn, src))
def par():
    k = 8; step = n // k
    th = [threading.Thread(target=lambda i=i: np.copyto(hn[i*step:(i+1)*step], src[i*step:(i+1)*step])) for i in range(k)]
    [t.start() for t in th]; [t.join() for t in th]
timed("refill 8 threads", par)
