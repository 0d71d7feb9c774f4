cd /tmp && python - <<'PY'
import numpy as np
g=np.load('/root/repo/tests/golden/ref_256_12289.npz')
np.savetxt('coeficientes_a.txt', g['fixture_a'][None], fmt='%d'); np.savetxt('coeficientes_b.txt', g['fixture_b'][None], fmt='%d')
PY
for i in 1 2 3; do /root/repo/ntt-based-polynomial-multiplier-fpga_b200/nttb200_time_testing256 | grep Tempo; done
