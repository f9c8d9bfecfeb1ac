import mpmath as mp, numpy as np
mp.mp.dps = 60
def fit(umax, n):
    # Chebyshev interpolation of g(u)=atan(sqrt(u))/sqrt(u) on [0,umax], n coefficients
    g = lambda u: mp.mpf(1) if u == 0 else mp.atan(mp.sqrt(u))/mp.sqrt(u)
    nodes = [mp.cos(mp.pi*(2*k+1)/(2*n)) for k in range(n)]
    us = [(x+1)/2*umax for x in nodes]
    # solve Vandermonde in mp for monomial coefs (n small)
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i,u in enumerate(us):
        for j in range(n): A[i,j] = u**j
        b[i] = g(u)
    c = mp.lu_solve(A, b)
    return [float(c[i]) for i in range(n)]
def test(c, umax, N=200001):
    r = np.linspace(-np.sqrt(umax), np.sqrt(umax), N)
    u = r*r
    p = np.full_like(u, c[-1])
    for k in range(len(c)-2, -1, -1): p = p*u + c[k]
    val = r*p
    ref = np.array([float(mp.atan(mp.mpf(x))) for x in r[::50]])
    err = np.abs(val[::50]-ref)/np.maximum(np.abs(ref),1e-300)
    return np.nanmax(err[np.abs(r[::50])>0])
for umax, ns in ((1.0, (18,20,22,24)), (float(mp.tan(mp.pi/8)**2), (10,11,12,13))):
    for n in ns:
        c = fit(umax, n)
        print(umax, n, test(c, umax))
