import ctypes, torch
n,k=4096,2048
A=torch.randn(k,n,dtype=torch.float64,device='cuda').t()   # n x k column-major, lda=n
C=torch.zeros(n,n,dtype=torch.float64,device='cuda').t()
cb=ctypes.CDLL("libcublas.so.12")
h=ctypes.c_void_p()
print('create',cb.cublasCreate_v2(ctypes.byref(h)))
one,zero=ctypes.c_double(1.0),ctypes.c_double(0.0)
st=cb.cublasDsyrk_v2(h,0,0,ctypes.c_int(n),ctypes.c_int(k),ctypes.byref(one),ctypes.c_void_p(A.data_ptr()),ctypes.c_int(A.stride(1)),ctypes.byref(zero),ctypes.c_void_p(C.data_ptr()),ctypes.c_int(C.stride(1)))
print('syrk status',st, A.stride(), C.stride())
torch.cuda.synchronize()
ref=(A@A.t()).tril()
print('maxdiff',(C.tril()-ref).abs().max().item())
