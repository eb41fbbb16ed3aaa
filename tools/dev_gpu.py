import sys, time, numpy as np
sys.path.insert(0,'.')
import rchem_b200 as rc
from rchem_b200 import geometry as geo
from oracle import oracle as orc
print("devices", rc.device_count())
z, x = geo.molecule(geo.WATER_CRAWFORD)
have_ref = orc.ref_lib() is not None
print("have _ref:", have_ref)
for name in ["STO-3G","6-31G","6-31G*"]:
    b = rc.Basis.new(z, x, name); ob = orc.make_basis(z, x, name)
    if have_ref: orc.use_reference_kernel(True)
    Iref = orc.build_I(ob)
    orc.use_reference_kernel(False) if have_ref else None
    Ix = orc.build_I(ob, orc.BOYS_EXACT)
    t=time.time(); I = rc.build_I(b); t1=time.time()-t
    print(name, "build_I ref-mode max diff vs ref: %.3e"%np.abs(I-Iref).max(), "time %.3f"%t1, b.stats())
    b.set_boys(1); I2 = rc.build_I(b)
    print(name, "build_I exact-mode vs exact oracle: %.3e   vs ref: %.3e"%(np.abs(I2-Ix).max(), np.abs(I2-Iref).max()))
    b.set_boys(0)
    n=b.nbf; D = geo.synthetic_density(n)
    J=np.zeros((n,n)); K=np.zeros((n,n)); rc.JK_direct(J,K,b,D)
    Jo,Ko = orc.jk_inmem(Iref, D)
    print(name, "JK_direct vs oracle JK_inmem(Iref): J %.3e K %.3e"%(np.abs(J-Jo).max(), np.abs(K-Ko).max()))
    Jg,Kg = rc.JK_inmem(I, D)
    print(name, "JK_inmem gpu vs oracle: J %.3e K %.3e"%(np.abs(Jg-Jo).max(), np.abs(Kg-Ko).max()))
# tier-1
ra=[1.,0.,1.]; rb=[0.,1.,2.]; rcc=[0.,0.,3.]; rd=[0.,0.,4.]
for pw, gold in [([0]*12, 0.08608517834596989), ([1]+[0]*11, -0.046888060557776973), ([2,1,0,1,0,0,1,0,0,0,1,0], 1.71817807954e-05)]:
    v = rc.coulomb_repulsion(1.1,1.2,1.3,1.4,ra,rb,rcc,rd,1.,1.,1.,1.,pw)
    vo = orc.coulomb_repulsion(ra,rb,rcc,rd,pw,[1.1,1.2,1.3,1.4],[1.]*4)
    print("tier1", pw, "gpu-oracle %.3e gpu-gold %.3e"%(v-vo, v-gold))
# bigger: (H2O)_3 6-31G* tensor + (H2O)_10 STO-3G JK timing
z3,x3 = geo.water_cluster(3); b=rc.Basis.new(z3,x3,"6-31G*"); ob=orc.make_basis(z3,x3,"6-31G*")
t=time.time(); I=rc.build_I(b); print("(H2O)3 6-31G* N=%d build_I %.3fs"%(b.nbf,time.time()-t), b.stats())
rng=np.random.default_rng(1); qs=rng.integers(0,b.nbf,size=(3000,4)).astype(np.int32)
if have_ref: orc.use_reference_kernel(True)
vals=orc.eval_quartets(ob, qs)
orc.use_reference_kernel(False) if have_ref else None
print("sampled 3000 integrals max diff %.3e"%np.abs(I[qs[:,0],qs[:,1],qs[:,2],qs[:,3]]-vals).max())
for nw,bas in [(10,"STO-3G"),(10,"6-31G*"),(32,"6-31G")]:
    zc,xc=geo.water_cluster(nw); b=rc.Basis.new(zc,xc,bas); n=b.nbf; D=geo.synthetic_density(n); J=np.zeros((n,n)); K=np.zeros((n,n))
    t=time.time(); rc.JK_direct(J,K,b,D); t0=time.time()-t
    t=time.time(); rc.JK_direct(J,K,b,D); t1=time.time()-t
    s=b.stats(); print("(H2O)%d %s N=%d JK first %.3fs second %.3fs"%(nw,bas,n,t0,t1), s, "GFLOP/s model %.1f"%(s['model_flops']/s['kernel_ms']/1e6))
