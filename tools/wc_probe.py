import ctypes, time, torch
rt = ctypes.CDLL('libcudart.so.12')
def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    return p.value
torch.cuda.init(); dev = torch.device('cuda:0')
N = 1 << 30
d = torch.empty(N, dtype=torch.uint8, device=dev)
d2 = torch.empty(N // 2, dtype=torch.uint8, device=dev)
st, st2 = torch.cuda.Stream(), torch.cuda.Stream()
def bw(ptr, label, with_d2h=None):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(N), 1, ctypes.c_void_p(st.cuda_stream))
        if with_d2h:
            rt.cudaMemcpyAsync(ctypes.c_void_p(with_d2h), ctypes.c_void_p(d2.data_ptr()), ctypes.c_size_t(N // 2), 2, ctypes.c_void_p(st2.cuda_stream))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(label, 'H2D %.1f GB/s' % (N / dt / 1e9), ('(+D2H %.1f GB/s)' % (N / 2 / dt / 1e9)) if with_d2h else '', flush=True)
plain = host_alloc(N, 0); wc = host_alloc(N, 4)   # cudaHostAllocWriteCombined = 4
out_plain = host_alloc(N // 2, 0)
ctypes.memset(plain, 1, N); ctypes.memset(wc, 1, N)
bw(plain, 'pinned      '); bw(wc, 'pinned WC   ')
bw(plain, 'pinned  +d2h', out_plain); bw(wc, 'pinned WC +d2h', out_plain)
