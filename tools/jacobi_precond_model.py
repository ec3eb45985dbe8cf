import numpy as np, sys, scipy.linalg as la
def rr_pairs(n, step):
    mm = n - 1
    pairs = [(mm, step % mm)]
    for k in range(1, n // 2):
        a = (step + k) % mm; b = (step - k) % mm
        pairs.append((min(a, b), max(a, b)))
    return pairs
def jac(X, b=8, order="blk", maxsw=40, eabs=float(sys.argv[2]) if len(sys.argv)>2 else 1e-15):
    X=X.copy(); n,L=X.shape; tol=np.sqrt(L)*2.2e-16; F=np.linalg.norm(X)
    nb=n//b
    if order=="blk":
        seq=[(I*b+p,I*b+q) for I in range(nb) for s in range(b-1) for (p,q) in rr_pairs(b,s)]+[(I*b+w,J*b+(w+s)%b) for r in range(nb-1) for (I,J) in rr_pairs(nb,r) for s in range(b) for w in range(b)]
    else: seq=[(p,q) for p in range(n) for q in range(p+1,n)]
    for sweep in range(maxsw):
        rot=0
        for p,q in seq:
            xp=X[p];xq=X[q]
            a=np.vdot(xp,xp).real;bb=np.vdot(xq,xq).real;c=np.dot(xp,xq.conj());ac=abs(c)
            if ac<=tol*np.sqrt(a*bb) or ac<=eabs*F*np.sqrt(max(a,bb)): continue
            rot+=1
            tau=(bb-a)/(2*ac); t=(1. if tau>=0 else -1.)/(abs(tau)+np.sqrt(1+tau*tau)); cs=1/np.sqrt(1+t*t); sn=t*cs; ph=c/ac
            yp=cs*xp-sn*ph*xq; yq=sn*np.conj(ph)*xp+cs*xq; X[p]=yp; X[q]=yq
        if rot==0: break
    return np.sort(np.linalg.norm(X,axis=1))[::-1], sweep+1
def mats(n):
    rs=np.random.RandomState(0)
    u,_=np.linalg.qr(rs.randn(n,n)+1j*rs.randn(n,n)); v,_=np.linalg.qr(rs.randn(n,n)+1j*rs.randn(n,n))
    yield "rand", rs.randn(n,n)+1j*rs.randn(n,n)
    yield "lowrank", (rs.randn(n,n//2)+1j*rs.randn(n,n//2))@(rs.randn(n//2,n)+1j*rs.randn(n//2,n))
    for d in (10,40,80): yield f"decay{d}", (u*np.exp(-np.arange(n)*d/n))@v
    D=np.exp(-np.arange(n)*30./n)
    yield "rowgraded", D[:,None]*(rs.randn(n,n)+1j*rs.randn(n,n))
    yield "colgraded", (rs.randn(n,n)+1j*rs.randn(n,n))*D[None,:]
    yield "colgraded_rev", (rs.randn(n,n)+1j*rs.randn(n,n))*D[None,::-1]
    yield "diagish", np.diag(D[::-1]).astype(complex)+1e-3*(rs.randn(n,n)+1j*rs.randn(n,n))*D[None,::-1]
n=int(sys.argv[1])
for name,m in mats(n):
    sref=np.linalg.svd(m,compute_uv=False)
    out=[]
    for pre in ("none","qr","qrcp","colsort_qr"):
        if pre=="none": X=m
        elif pre=="qr": X=np.linalg.qr(m)[1]
        elif pre=="qrcp": X=la.qr(m,pivoting=True)[1]
        else:
            o=np.argsort(-np.linalg.norm(m,axis=0)); X=np.linalg.qr(m[:,o])[1]
        s,nsw=jac(X)
        out.append(f"{pre}:{nsw}sw err{np.max(np.abs(s-sref))/sref[0]:.0e}")
    print(name, " | ".join(out))
