cd /root/repo
python - <<'PY'
import numpy as np, sys
sys.path.insert(0,'.')
from rududu_image_codec_b200.synth import synth_image
img = synth_image(3, 250, 134, 3)
with open('/tmp/a.ppm','wb') as f:
    f.write(b"P6\n250 134\n255\n"); f.write(np.ascontiguousarray(img.transpose(1,2,0)).tobytes())
PY
for env in "RIC_FWD0=1 RIC_QUANT_PK=1" "RIC_FWD0=0 RIC_QUANT_PK=1" "RIC_FWD0=1 RIC_QUANT_PK=0" "RIC_FWD0=0 RIC_QUANT_PK=0"; do
  echo "== $env"; env $env timeout 20 rududu_image_codec_b200/ric_b200 -i /tmp/a.ppm -o /tmp/a.ric -q 9; echo "rc=$?"; ls -la /tmp/a.ric 2>/dev/null; rm -f /tmp/a.ric
done
