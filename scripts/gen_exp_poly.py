"""Coefficients of the degree-N polynomial used by the in-kernel exp of the softmax operator (operators.cuh).

exp(r) on |r| <= ln2/2 by interpolation at Chebyshev nodes in 60-digit arithmetic (near-minimax), coefficients
rounded to double; prints the max relative error of the *rounded* polynomial evaluated exactly.
  python scripts/gen_exp_poly.py [degree] [halfwidth divisor: 2 -> |r| <= ln2/2 (default), 64 -> |r| <= ln2/64 + table]
With a divisor > 2 the 2^(j/32) table of the table-driven variant (exp_bounded_tab) is printed as well.
"""
import struct
import sys

import mpmath as mp

mp.mp.dps = 60
deg = int(sys.argv[1]) if len(sys.argv) > 1 else 10
div = int(sys.argv[2]) if len(sys.argv) > 2 else 2
h = mp.log(2) / div * mp.mpf("1.0001")
n = deg + 1
nodes = [h * mp.cos(mp.pi * (2 * i + 1) / (2 * n)) for i in range(n)]
V = mp.matrix(n, n)
for i, x in enumerate(nodes):
    for j in range(n):
        V[i, j] = x ** j
c = mp.lu_solve(V, mp.matrix([mp.exp(x) for x in nodes]))
cd = [float(c[j]) for j in range(n)]
cd[0] = 1.0
worst = mp.mpf(0)
for i in range(4001):
    x = -h + 2 * h * i / 4000
    p = sum(mp.mpf(cd[j]) * x ** j for j in range(n))
    worst = max(worst, abs(p / mp.exp(x) - 1))
print(f"// degree {deg}, max relative error of the rounded polynomial on |r| <= ln2/{div}: {mp.nstr(worst, 3)}")
for j in range(n):
    print(f"    {cd[j]!r},  // 0x{struct.unpack('<Q', struct.pack('<d', cd[j]))[0]:016x}")
if div > 2:
    m = div // 2
    print(f"// 2^(j/{m}), j = 0..{m - 1}")
    for j in range(m):
        v = float(mp.mpf(2) ** (mp.mpf(j) / m))
        print(f"    {v!r},  // 0x{struct.unpack('<Q', struct.pack('<d', v))[0]:016x}")
