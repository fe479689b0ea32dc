cd scripts
for pw in 1 2 4 8 16 99; do
 echo "prewarps=$pw: $(RMT_EXT_PREWARPS=$pw python -c "
from extrap_probe import case
g=lambda k:[((m+.5)/k,(n+.5)/k) for n in range(k) for m in range(k)]
print('64 discs %.3f | 1 disc %.3f | 16 discs %.3f' % (case(4097,g(8),164), case(4097,[(0.5,0.5)],164), case(4097,g(4),164)))")"
done
