import sys, ctypes as C, numpy as np, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gardenia_b200 as gb
L = C.CDLL('' + os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libband.so') + '')
kind, scale = sys.argv[1], int(sys.argv[2])
t=time.time(); g = gb.Graph.generate(kind, scale, 16); print('gen', time.time()-t, g.m, g.nnz, flush=True)
rp, ci = g.out_rowptr(), g.out_colidx()
deg = np.diff(rp).astype(np.int64)
perm = np.argsort(-deg, kind='stable').astype(np.int32)
newid = np.empty(g.m, np.int32); newid[perm] = np.arange(g.m, dtype=np.int32)
n_nz = int((deg>0).sum())
print('nonzero rows', n_nz, 'rows deg>=64', int((deg>=64).sum()), '>=256', int((deg>=256).sum()), '>=1024', int((deg>=1024).sum()))
sd = deg[perm]
cs = np.cumsum(sd)
for k in (49152, 1<<20, 4<<20): 
    if k < g.m: print('edge share ids<%d: %.3f' % (k, cs[k-1]/g.nnz))
for th in (64,128,256,512,1024): print('edges in rows deg>=%d: %.3f' % (th, sd[sd>=th].sum()/g.nnz))
band = 49152
for B, cmin in [(int(a), int(b)) for a, b in (x.split(',') for x in sys.argv[3:])]:
    out = np.zeros(B*4, np.int64); mp = np.zeros(2, np.int64)
    t=time.time()
    L.band_stats(C.c_int64(g.m), rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p), newid.ctypes.data_as(C.c_void_p), perm.ctypes.data_as(C.c_void_p), B, band, cmin, out.ctypes.data_as(C.c_void_p), mp.ctypes.data_as(C.c_void_p))
    o = out.reshape(B,4)
    print(f'B={B} cmin={cmin}: pairs={o[:,0].sum()/1e6:.1f}M moved={o[:,1].sum()/g.nnz:.3f} of nnz (band0 {o[0,1]/g.nnz:.3f}) padded_u16={o[:,2].sum()/1e6:.0f}M ({o[:,2].sum()/max(1,o[:,1].sum()):.2f}x) nosort={o[:,3].sum()/max(1,o[:,1].sum()):.2f}x  main remaining={mp[0]/g.nnz:.3f} padded {mp[1]/max(1,mp[0]):.3f}x  ({time.time()-t:.0f}s)', flush=True)
    if '-v' in sys.argv: print(o[:8])
