#!/bin/bash
cd "$(dirname "$0")"
for cfg in "256 32 20" "256 32 0" "512 32 0" "128 32 0" "256 16 0" "256 64 0" "1024 32 0"; do
  set -- $cfg
  echo "XT=$1 MRB=$2 SLEEP=$3: $(RMT_EXT_XT=$1 RMT_EXT_MRB=$2 RMT_EXT_SLEEP=$3 python -c "
from extrap_probe import case
g=lambda k:[((m+.5)/k,(n+.5)/k) for n in range(k) for m in range(k)]
print('64 discs %.3f ms | 1 disc %.3f ms' % (case(4097,g(8),164), case(4097,[(0.5,0.5)],164)))")"
done
