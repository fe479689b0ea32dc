#!/bin/bash
# extrapolation time at 4097^2 for 64 / 1 / 16 / 4 discs, all-layers kernel allowed vs per-layer launches only
cd "$(dirname "$0")"
for f in 1 0; do
  echo "FUSED=$f: $(RMT_EXT_FUSED=$f python -c "
from extrap_probe import case
g=lambda k:[((m+.5)/k,(n+.5)/k) for n in range(k) for m in range(k)]
print('64 discs %.3f ms | 1 disc %.3f ms | 16 discs(4x4,R164) %.3f | 36 discs(6x6) %.3f | 4 discs R400 %.3f' % (case(4097,g(8),164), case(4097,[(0.5,0.5)],164), case(4097,g(4),164), case(4097,g(6),164), case(4097,g(2),400)))")"
done
