"""throughput of the device log-mel front end (rnnt/features.py + transforms.Downsample restated in
csrc/frontend.cu + eb_gemm_f32): B utterances of 16 s at 16 kHz -> [B, T, 240] model input (80 mel x 3 stacked frames),
audio-seconds per second, with the SpecAugment masks on top."""
import json, os, sys, random
import torch
sys.path.insert(0, os.getcwd())
from edgedict_b200.rnnt.features import LogMelFrontend, TimeMasking, FrequencyMasking
B, SEC = 32, 16
x = torch.randn(B, SEC * 16000, device="cuda") * 0.1
fe = LogMelFrontend(feature_size=80, n_fft=512, win_length=400, hop_length=200, downsample=3, dither=0.0).cuda()
tm, fm = TimeMasking(50, 2), FrequencyMasking(27, 2)


def run(mask):
    y = fe(x)
    if mask:
        y = fm(tm(y.transpose(1, 2))).transpose(1, 2)
    return y


for mask in (False, True):
    for _ in range(3):
        y = run(mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        y = run(mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps(dict(frontend="log-mel 80 x 3, n_fft 512, win 400, hop 200", specaugment=mask, batch=B, seconds_each=SEC,
                          out_shape=list(y.shape), ms=round(ms, 3), audio_sec_per_sec=round(B * SEC / ms * 1e3, 1))), flush=True)
