import ctypes, sys, torch
sys.path.insert(0, '.')
from last_torch_b200 import _native as N
N.lib()
h = ctypes.CDLL(N.LIB_PATH)
fn = h.ltx_umma_probe_mn
fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
swap = int(sys.argv[1]); n = int(sys.argv[2]); k = int(sys.argv[3])
at = torch.randn([k, 128], device='cuda'); bt = torch.randn([k, n], device='cuda')
d = torch.zeros([128, n], device='cuda')
rc = fn(at.data_ptr(), bt.data_ptr(), d.data_ptr(), n, k, swap, None)
torch.cuda.synchronize()
ref = at.double().T @ bt.double()
print('swap', swap, 'n', n, 'k', k, 'rc', rc, 'relerr', float((d.double() - ref).abs().max() / ref.abs().max()))
