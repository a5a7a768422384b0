import numpy as np
from scipy.special import erf
np.set_printoptions(precision=17)
SQPI = np.sqrt(np.pi)
def V(w):
    w = np.asarray(w, dtype=np.float64); z = np.sqrt(np.maximum(w, 1e-300))
    small = w < 1e-2
    # series erf(z)/z = 2/sqrt(pi) sum (-1)^n w^n / (n! (2n+1))
    s = np.zeros_like(w); term = np.ones_like(w)
    for n in range(0, 12):
        s += term / (2*n+1); term = term * (-w) / (n+1)
    out = np.where(small, 2/SQPI*s, erf(z)/z)
    return out
def B(w):
    w = np.asarray(w, dtype=np.float64); z = np.sqrt(np.maximum(w, 1e-300))
    small = w < 0.5
    # B = 2 dV/dw = 2/sqrt(pi) * 2 * sum_{n>=1} (-1)^n n w^(n-1) / (n! (2n+1))
    s = np.zeros_like(w); 
    from math import factorial
    for n in range(1, 30):
        s += ((-1)**n) * n * w**(n-1) / (factorial(n) * (2*n+1))
    ser = 2/SQPI*2*s
    with np.errstate(all='ignore'):
        closed = 2/SQPI*np.exp(-w)/w - erf(z)/z**3
    return np.where(small, ser, closed)

def ratfit(f, W, np_, nq, iters=60, npts=4000):
    # chebyshev-distributed nodes on [0, W]
    k = np.arange(npts); u = 0.5*(1-np.cos(np.pi*(k+0.5)/npts))  # in [0,1]
    x = u*W; fx = f(x)
    wt = np.ones(npts); Qprev = np.ones(npts)
    best=None
    for it in range(iters):
        # unknowns p0..p_np, q1..q_nq  in variable u
        A = np.zeros((npts, np_+1+nq))
        for i in range(np_+1): A[:, i] = u**i
        for j in range(1, nq+1): A[:, np_+j] = -fx*u**j
        rhs = fx.copy()
        scale = wt/(np.abs(fx)*np.abs(Qprev))
        sol, *_ = np.linalg.lstsq(A*scale[:,None], rhs*scale, rcond=None)
        p = sol[:np_+1]; q = np.concatenate([[1.0], sol[np_+1:]])
        P = sum(p[i]*u**i for i in range(np_+1)); Q = sum(q[j]*u**j for j in range(nq+1))
        err = (P/Q - fx)/fx
        m = np.max(np.abs(err))
        if best is None or m < best[0]: best=(m, p.copy(), q.copy())
        # Lawson reweighting
        wt = wt*(0.2+np.abs(err)/m); wt /= wt.mean()
        Qprev = Q
    m,p,q = best
    # convert to coefficients in w: u = w/W
    pw = np.array([p[i]/W**i for i in range(np_+1)]); qw = np.array([q[j]/W**j for j in range(nq+1)])
    return m, pw, qw

def eval32(pw, qw, w):
    w = w.astype(np.float32)
    P = np.float32(pw[-1])*np.ones_like(w)
    for c in pw[-2::-1]: P = P*w + np.float32(c)
    Q = np.float32(qw[-1])*np.ones_like(w)
    for c in qw[-2::-1]: Q = Q*w + np.float32(c)
    return P/Q

import sys
W = float(sys.argv[1]) if len(sys.argv)>1 else 16.0
for name,f in (("V",V),("B",B)):
    for (a,b) in ((5,4),(6,4),(6,5),(7,5),(6,6),(7,6)):
        m,pw,qw = ratfit(f, W, a, b)
        x = np.linspace(0, W, 200001)
        e32 = np.max(np.abs((eval32(pw,qw,x).astype(np.float64)-f(x))/f(x)))
        print(name, (a,b), "fit max rel %.2e  float32 eval max rel %.2e  minQ %.3g"%(m, e32, np.min(np.polyval(qw[::-1], x))))

def eval_fma(pw, qw, w):
    w = w.astype(np.float32)
    def horner(c):
        acc = np.full_like(w, np.float32(c[-1]))
        for k in c[-2::-1]:
            acc = (acc.astype(np.float64)*w.astype(np.float64) + np.float64(np.float32(k))).astype(np.float32)
        return acc
    P = horner(pw); Q = horner(qw)
    return (P.astype(np.float64)/Q.astype(np.float64)).astype(np.float32)
print("---- chosen")
for name,f,(a,b) in (("V",V,(6,4)),("B",B,(6,5)),("V",V,(6,5))):
    m,pw,qw = ratfit(f, W, a, b, iters=120)
    x = np.linspace(0, W, 400001)
    e = np.abs((eval_fma(pw,qw,x).astype(np.float64)-f(x))/f(x))
    print(name,(a,b),"fit %.2e fma-eval max rel %.2e rms %.2e"%(m, e.max(), np.sqrt((e**2).mean())))
    print(" P:", ", ".join("%.9ef"%c for c in pw))
    print(" Q:", ", ".join("%.9ef"%c for c in qw))
