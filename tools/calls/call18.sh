set -x
timeout 900 python -m pytest tests/test_image_encoder_gpu.py tests/test_pipeline_gpu.py tests/test_preprocess_gpu.py -m gpu -q -s 2>&1 | grep -E "DINO|passed|failed|FAILED|Error|assert" | cut -c1-300
timeout 300 python - <<'PY'
import time, torch
from actionmesh_b200.image_encoder import B200ImageEncoder
for prec in ("fp32", "bf16"):
    enc = B200ImageEncoder(precision=prec).to("cuda"); enc.init_random_()
    px = torch.randn(16, 3, 224, 224).cuda()
    for _ in range(2): enc.encode_pixel_values(px)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(3): enc.encode_pixel_values(px)
    torch.cuda.synchronize(); print("DINO_ENCODE_16_FRAMES_S", prec, (time.time() - t0) / 3)
PY
