"""Checks the `triple a b r` lines of tools/mulbench5 (the 29-bit candidate's product): r == a b 2^-261 mod p."""
import sys
P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
ok = n = 0
for line in open(sys.argv[1]):
    if line.startswith("triple"):
        a, b, r = (int(x, 16) for x in line.split()[1:])
        n += 1
        ok += (r % P) == (a * b * pow(1 << 261, -1, P)) % P
print(f"{ok} of {n} products correct")
sys.exit(0 if ok == n and n else 1)
