"""Numerics experiment behind profiles/r2_tcgen05_probe.md (tool, numpy only): the interior-point method of the oracle on the
joint QP of BASELINE configs[1] with the reduced Hessian factorised in float32 (what a 3 x TF32 tcgen05 factorisation delivers)
after Jacobi scaling, plus k steps of FP64 iterative refinement of every Newton solve."""
import sys, numpy as np, scipy.linalg as sl
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import oracle, oracle_util, feas_util as fu
def pdip_mixed(Q,A,b,G,h, fdt, refine, tol_gap=1e-10, tol_res=1e-9, max_iter=60, verbose=False):
    P = Q+Q.T
    xp = np.linalg.lstsq(A,b,rcond=None)[0]; Z = sl.null_space(A)
    GZ = G@Z; live = np.abs(GZ).max(axis=1) > 1e-12
    G=G[live]; h=h[live]; mi=len(h)
    stats=[]
    def mk(w):
        H = Z.T@(P + G.T@(w[:,None]*G))@Z
        # symmetric diagonal scaling (Jacobi) before the low-precision factorisation
        dsc = 1/np.sqrt(np.diag(H)); Hs = (H*dsc[:,None])*dsc[None,:]
        try: c = sl.cho_factor(Hs.astype(fdt))
        except Exception: return None
        def solve(rhs):
            r = Z.T@rhs
            def s32(v): return dsc*sl.cho_solve(c,(dsc*v).astype(fdt)).astype(np.float64)
            d = s32(r); rn0=np.abs(r).max()
            for k in range(refine):
                res = r - H@d
                d = d + s32(res)
            res = r - H@d
            stats.append(np.abs(res).max()/max(rn0,1e-300))
            return d
        return solve
    x = xp.copy()
    dsg = mk(np.ones(mi))(-P@x + G.T@(h-G@x)); x = x+Z@dsg
    z = G@x-h; s=-z; ap=(-s).max(); ad=(-z).max()
    if ap>=0: s+=1+ap
    if ad>=0: z+=1+ad
    hn=np.abs(h).max()
    for it in range(max_iter):
        px=P@x; rd = px+G.T@z; rg=G@x+s-h; mu=s@z/mi
        nrd=np.abs(Z.T@rd).max(); nrg=np.abs(rg).max(); obj=0.5*x@px; dsc_=1+np.abs(px).max()
        if verbose: print(it,'mu %.1e nrd %.1e wmax %.1e linres %s'%(mu,nrd,(z/s).max(), ' '.join('%.0e'%v for v in stats[-2:])))
        gap_ok = mu<=tol_gap*max(1,abs(obj)) and nrg<=tol_res*(1+hn)
        if gap_ok and nrd<=1e-6*dsc_: return 0,it,x
        w=z/s
        solve=mk(w)
        if solve is None: return 2,it,x
        dsa = solve(-rd + G.T@(-(w*rg-z)))
        dxa=Z@dsa; gxa=G@dxa
        dsa_=-rg-gxa; dza=-z-w*dsa_
        aa=1.0
        for v,d in ((s,dsa_),(z,dza)):
            neg=d<0
            if neg.any(): aa=min(aa,(-v[neg]/d[neg]).min())
        mua=((s+aa*dsa_)*(z+aa*dza)).sum()/mi
        sigma=(mua/mu)**3
        rc=s*z+dsa_*dza-sigma*mu
        tt=-(z*rg-rc)/s
        dsg=solve(-rd+G.T@tt); dx=Z@dsg; gx=G@dx
        ds=-rg-gx; dz=(-rc-z*ds)/s
        am=1e300
        for v,d in ((s,ds),(z,dz)):
            neg=d<0
            if neg.any(): am=min(am,(-v[neg]/d[neg]).min())
        al=min(1,0.99*am)
        x+=al*dx; s+=al*ds; z+=al*dz
    return 2,it,x
ms=fu.missions('cfg2',6)
for c,m in enumerate(ms[:6]):
    p=oracle_util.oracle_problem(m, sequential=False, batch_size=16)
    q=p.populate(np.zeros((16*30,3)),0)
    Q,A,b,G,h=q.dense()
    r64=pdip_mixed(Q,A,b,G,h,np.float64,0)
    out=[]
    for refine in (0,1,2,4):
        r=pdip_mixed(Q,A,b,G,h,np.float32,refine)
        out.append((refine,r[0],r[1],'%.1e'%np.abs(r[2]-r64[2]).max()))
    print('mission',c,'fp64:',r64[:2],'| fp32 factor (refine steps, status, iters, |dx|):',out,flush=True)
