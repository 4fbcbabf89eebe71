import sys, numpy as np, torch
sys.path.insert(0, ".")
from ser_b200 import _native, synth
from ser_b200.config import FeatureFlags, flag_bits
SR = 48000
ctx = _native.get_context(0); bits = flag_bits(FeatureFlags())
for seconds, batch in ((60, 512), (60, 1), (3.5, 4096)):
    n = int(seconds * SR)
    base = synth.batch_audio_torch(min(batch, 64), SR, n, device="cuda")
    wave = base.repeat((batch + 63) // 64, 1)[:batch].contiguous().reshape(-1)
    starts = np.arange(batch, dtype=np.int64) * n; lengths = np.full(batch, n, dtype=np.int64)
    out = torch.empty((batch, 193), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    for _ in range(2): ctx.features_device(wave.data_ptr(), wave.numel(), starts, lengths, SR, bits, out.data_ptr(), 0)
    torch.cuda.synchronize()
    ctx.set_profile(True)
    ctx.features_device(wave.data_ptr(), wave.numel(), starts, lengths, SR, bits, out.data_ptr(), 0)
    torch.cuda.synchronize()
    k = ctx.kernel_ms(); ctx.set_profile(False)
    print(seconds, batch, {a: round(b[0], 2) for a, b in k.items() if b[0] > 0})
    del wave, base
