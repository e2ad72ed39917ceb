import ctypes, os, torch
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "wbw.so"))
lib.wbw.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
n = 8 * 1024 ** 3
buf = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
def t(variant, blocks):
    for _ in range(2): lib.wbw(buf.data_ptr(), n, variant, blocks, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): rc = lib.wbw(buf.data_ptr(), n, variant, blocks, s)
    e1.record(); torch.cuda.synchronize()
    return n / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e12, rc
names = {0: "st.v4", 1: "st.cs.v4", 2: "st.L1::no_allocate.v4", 3: "st.wt.v4", 4: "st.v2 (8 B)", 5: "cudaMemsetAsync"}
for v in range(6):
    for blocks in ((148 * 8, 148 * 32, 148 * 128) if v != 5 else (1,)):
        bw, rc = t(v, blocks)
        print(f"{names[v]:24s} blocks {blocks:6d}: {bw:.2f} TB/s rc {rc}")
