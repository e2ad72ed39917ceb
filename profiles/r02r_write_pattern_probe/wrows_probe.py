import ctypes, os, torch
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "wbw.so"))
lib.wrows.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
G = 4096
buf = torch.empty(G * 1000 * 2048, dtype=torch.uint8, device="cuda")
sink = torch.zeros(1, device="cuda")
s = torch.cuda.current_stream().cuda_stream
def t(row_bytes, v16, work, blocks):
    for _ in range(2): lib.wrows(buf.data_ptr(), G, row_bytes, v16, work, blocks, sink.data_ptr(), s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): rc = lib.wrows(buf.data_ptr(), G, row_bytes, v16, work, blocks, sink.data_ptr(), s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    return ms, G * 1000 * row_bytes / (ms * 1e-3) / 1e12
for row_bytes in (2000, 2048):
    for v16 in (0, 1):
        for work in (0, 200):
            for blocks in (592, 1184):
                ms, bw = t(row_bytes, v16, work, blocks)
                print(f"row {row_bytes} B  {'16' if v16 else ' 8'}-byte stores  work {work:3d}  blocks {blocks:5d}: {ms:.3f} ms  {bw:.2f} TB/s")
