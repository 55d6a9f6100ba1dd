import numpy as np
from scipy.optimize import linprog
def minimax(deg, B, n=4001):
    x = np.cos(np.pi*(np.arange(n)+0.5)/n)*B   # chebyshev-dense grid incl. near the ends
    x = np.concatenate([x, [-B, B]])
    f = 2.0**x
    # variables: c0..cdeg, e ; constraints |sum c x^k / f - 1| <= e
    V = np.vander(x, deg+1, increasing=True)/f[:,None]
    A = np.block([[V, -np.ones((len(x),1))],[-V, -np.ones((len(x),1))]])
    b = np.concatenate([np.ones(len(x)), -np.ones(len(x))])
    c = np.zeros(deg+2); c[-1]=1
    r = linprog(c, A_ub=A, b_ub=b, bounds=[(None,None)]*(deg+1)+[(0,None)], method="highs")
    return r.x[:-1], r.x[-1]
for deg,B in [(2,0.5),(2,0.75),(3,1.0),(3,1.25),(4,1.0),(4,1.5),(4,2.0),(4,2.5),(5,2.0),(5,3.0),(6,3.0),(6,4.0)]:
    c,e = minimax(deg,B)
    print(deg,B,f"err={e:.3e}", ", ".join(f"{v:.9g}" for v in c))
print("----")
for deg,B in [(3,0.5),(3,0.75),(4,1.25),(5,2.5),(5,2.75),(6,3.25),(6,3.5)]:
    c,e = minimax(deg,B)
    print(deg,B,f"err={e:.3e}", ", ".join(f"{v:.9g}f" for v in c))
