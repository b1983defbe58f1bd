"""Measured HBM throughput for independent random reads of 32/64/128-byte chunks (B200),
as a function of the footprint (TLB reach is ~256 MB: SURVEY / B300_MICROARCH)."""
import ctypes as C
import sys
sys.path.insert(0, ".")
from alphazero_othello_b200 import _lib
L = _lib.lib()
for mib in (512, 2048, 8192, 24576, 40960):
    row = []
    for chunk in (32, 64, 128):
        g, ms = C.c_double(0), C.c_float(0)
        _lib.check(L.oth_host_random_read_probe(mib << 20, chunk, C.byref(g), C.byref(ms)))
        row.append(f"{chunk}B {g.value:7.1f} GB/s")
    print(f"footprint {mib:6d} MiB: " + "  ".join(row))
