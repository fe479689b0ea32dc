#!/bin/bash
# extrapolation time at 4097^2 for several body counts: automatic choice, per-layer launches only, and each
# all-layers variant forced (RMT_EXT_FORCE = rows-per-block * 10000 + macro-tile rows)
cd "$(dirname "$0")"
run() {
  python -c "
from extrap_probe import case
g=lambda k:[((m+.5)/k,(n+.5)/k) for n in range(k) for m in range(k)]
print('64 discs %.3f ms | 1 disc %.3f ms | 16 discs(4x4,R164) %.3f | 36 discs(6x6) %.3f | 4 discs R400 %.3f' % (case(4097,g(8),164), case(4097,[(0.5,0.5)],164), case(4097,g(4),164), case(4097,g(6),164), case(4097,g(2),400)))"
}
echo "auto        : $(run)"
echo "per-layer   : $(RMT_EXT_FUSED=0 run)"
for f in 160512 161024 80512 81024; do echo "force $f: $(RMT_EXT_FORCE=$f run)"; done
