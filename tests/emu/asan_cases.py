import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, ctypes
from tests.emu import emu
emu.LIB=os.environ.get('PCL_EMU_LIB','/tmp/libpcl_emu_asan.so')
emu._lib=ctypes.CDLL(emu.LIB); emu._lib.pcl_last_error.restype=ctypes.c_char_p; emu._lib.simt_set_reverse.argtypes=[ctypes.c_int]
from oracle import oracle
import polarcode_and_ldpc_b200 as P
rng=np.random.default_rng(0)
for N,K,L,F in ((256,128,8,5),(64,32,32,3),(1024,512,8,2),(16,8,4,9),(128,64,2,19)):
    fz=P.bhattacharyya_frozen_set(N,K,2.0); llr=rng.normal(1,3,size=(F,N))
    ref=oracle.polar_scl(N,L,fz,llr)
    for S in (1,0):
        for dt in ('f32','f64'):
            got=emu.polar_decode(N,K,L,fz,llr,dt,want_pm=True,want_leaf=True,env={'PCL_POLAR_NL':S})[0]
            assert np.array_equal(got,ref),(N,L,S,dt)
    got=emu.polar_decode(N,K,L,fz,llr,'f64',crc=(0x1D,8)); 
    got=emu.polar_decode(N,K,L,fz,llr,'f64',env={'PCL_POLAR_GENERIC':1}); assert np.array_equal(got,ref)
for n in (96,504):
    H=P.gallager_parity_check(n,3,6,42); llr=rng.normal(1,2.5,size=(4,n))
    for mode in ('bp','ms'):
        for dt in ('f32','f64'):
            emu.ldpc_decode(H,llr,mode,6,0.75,True,dt,want_total=True)
Hm=P.mackay_parity_check(120,60,3,6,seed=42); emu.ldpc_decode(Hm,rng.normal(1,2,size=(3,120)),'bp',5)
# block-per-frame mode and the check-major layout in the fp32 build
for env in ({'PCL_LDPC_COOP':'1'},{'PCL_LDPC_BANKED':'0'},{'PCL_LDPC_COOP':'1','PCL_LDPC_BANKED':'0'}):
    os.environ.update(env)
    H=P.gallager_parity_check(96,3,6,42)
    for mode in ('bp','ms'):
        emu.ldpc_decode(H,rng.normal(1,2.5,size=(3,96)),mode,4,0.75,True,'f32',want_total=True)
    for k in env: os.environ.pop(k)
# frame generator: polar / LDPC, every channel, odd sizes
for N,K in ((16,7),(64,33),(1024,512),(2048,1000)):
    fz=P.bhattacharyya_frozen_set(N,K,2.0)
    for ch in (0,1,2):
        emu.gen_frames('polar',N,K,fz,5,0.1 if ch==2 else 1.5,seed=3,frame0=7,channel=ch,dtype='f64' if ch==1 else 'f32')
for n in (96,504):
    H=P.gallager_parity_check(n,3,6,42); G,_=P.generator_from_parity(H)
    emu.gen_frames('ldpc',n,G.shape[0],G,5,1.0,seed=4)
print('asan run ok')
# round 2: lists wider than a warp (block per frame), checks wider than 32 edges, tensor-memory variants of other sizes
g=np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'golden','wide.npz'))
for ci in range(int(g['nscl'])):
    N,L,fz,llr=int(g[f'scl{ci}_N']),int(g[f'scl{ci}_L']),g[f'scl{ci}_frozen'],g[f'scl{ci}_llr']
    for dt in ('f32','f64'):
        got=emu.polar_decode(N,N-len(fz),L,fz,llr,dt,want_pm=True,want_leaf=(L<=256))[0]
        if dt=='f64': assert np.array_equal(got,g[f'scl{ci}_bits']),ci
emu.polar_decode(128,70,64,P.bhattacharyya_frozen_set(128,70,2.0),rng.normal(1,3,size=(2,128)),'f32',crc=(0x1D,8))
for dt in ('f32','f64'):
    emu.ldpc_decode(g['dense_bp_H'].astype(np.int64),g['dense_bp_llr'],'bp',6,1.0,True,dt,want_total=True)
for N,K,L,F in ((1024,512,16,3),(1024,512,32,2),(2048,1024,8,5)):
    fz=P.bhattacharyya_frozen_set(N,K,2.0); llr=rng.normal(1,3,size=(F,N))
    assert np.array_equal(emu.polar_decode(N,K,L,fz,llr,'f32'),oracle.polar_scl(N,L,fz,llr)),(N,L)
fz=P.bhattacharyya_frozen_set(1024,512,2.0); llr=rng.normal(1,3,size=(37,1024))
assert np.array_equal(emu.polar_decode(1024,512,1,fz,llr,'f32'),oracle.polar_sc(1024,fz,llr))     # polar_sc_big_kernel<4>
for N2 in (512, 2048):
    fz=P.bhattacharyya_frozen_set(N2,N2//2,2.0); llr=rng.normal(1,3,size=(33,N2))
    assert np.array_equal(emu.polar_decode(N2,N2//2,1,fz,llr,'f32'),oracle.polar_sc(N2,fz,llr))   # polar_sc_big_kernel<2>, <8>
print('asan round-2 cases ok')
