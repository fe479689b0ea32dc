#!/usr/bin/env python
"""Regenerate exp_table.inc: the 2^(k/128) table of the table-driven exp() that
glibc >= 2.28 uses (Szabolcs Nagy / ARM optimized-routines, EXP_TABLE_BITS = 7).

For k in [0, 128):  H = RN(2^(k/128)),  tail = RN(2^(k/128)/H - 1),
entry = (bits(tail), bits(H) - (k << 52)/128).   Computed with 80-digit decimals.
extrap.cu evaluates exp() with this table and explicit FMAs so the device result
is bit-identical to the host libm exp() that the reference (Numba) calls for the
least-squares weights (upstream pyRMT/functions.py:120).
"""
import os
import struct
from decimal import Decimal, getcontext

getcontext().prec = 80
LN2 = Decimal(2).ln()


def bits(x):
    return struct.unpack('<Q', struct.pack('<d', x))[0]


def main():
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "exp_table.inc")
    with open(out, "w") as f:
        for k in range(128):
            v = (LN2 * k / 128).exp()
            H = float(v)
            tail = float(v / Decimal(H) - 1)
            f.write("0x%016xULL, 0x%016xULL,\n"
                    % (bits(tail), (bits(H) - ((k << 52) // 128)) & 0xFFFFFFFFFFFFFFFF))


if __name__ == "__main__":
    main()
