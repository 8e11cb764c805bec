"""Cluster/tcgen05 recurrent kernels (csrc/lstm_c4.cu) against the mma.sync kernels (csrc/lstm_tc.cu) on the same
inputs, then timing alone / paired on two streams.  usage: gpu_c4_check.py [check] [time]"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from edgedict_b200 import ops
from edgedict_b200._lib import lib

dev = "cuda"
what = sys.argv[1:] or ["check", "time"]


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def unpack_saves(gs, cs, B, T, H):
    """CTA-private saves -> gates [B,T,4H], c [B,T,H] (first batch tile only)."""
    NC = H // 8
    g = gs.view(torch.bfloat16).view(-1, T, NC, 32, 4, 4, 2)[0].float()       # [T][cta][b][up][gate][e]
    gates = g.permute(2, 0, 4, 1, 3, 5).reshape(32, T, 4, H)                   # b, T, gate, (cta, up, e) = unit
    c = cs.view(torch.float32).view(-1, T, NC, 32, 4, 2)[0].permute(2, 0, 1, 3, 4).reshape(32, T, H)
    return gates.reshape(32, T, 4 * H)[:B], c[:B]


if "check" in what:
    print("c4 supported:", {H: lib().eb_lstm_c4_supported(32, H) for H in (256, 512, 768, 1024)},
          "bwd cluster:", {H: lib().eb_lstm_c4_bwd_cluster(H) for H in (256, 512, 768, 1024)}, flush=True)
    for (B, T, H) in [(32, 6, 256), (5, 9, 512), (32, 37, 1024), (40, 5, 768), (32, 300, 1024)]:
        torch.manual_seed(B * 1000 + T)
        xg = torch.randn(B, T, 4 * H, device=dev)
        whh = (torch.rand(4 * H, H, device=dev) * 2 - 1) / H ** 0.5
        whh16 = whh.bfloat16()
        whhT16 = whh16.t().contiguous()
        h0, c0 = torch.randn(B, H, device=dev) * 0.5, torch.randn(B, H, device=dev) * 0.5
        y0, y16, hT0, cT0, gates0, cseq0 = ops.lstm_tc_fwd(xg, whh16, h0, c0, True)
        y1, hp1, hT1, cT1, gs, cs = ops.lstm_c4_fwd(xg, whh16, h0, c0, True)
        gstd, cstd = ops.lstm_c4_fwd(xg, whh16, h0, c0, True, std_saves=True)[4:]
        print("    std saves: gates %.2e c %.2e" % (rel(gstd, gates0), rel(cstd, cseq0)))
        torch.cuda.synchronize()
        hp_ref = torch.cat([h0.bfloat16()[:, None], y16[:, :-1]], 1)
        print("fwd B%d T%d H%d: y %.2e hT %.2e cT %.2e hprev %.2e" % (B, T, H, rel(y1, y0), rel(hT1, hT0), rel(cT1, cT0),
                                                                      rel(hp1.float(), hp_ref.float())), flush=True)
        if B <= 32:
            g1, c1 = unpack_saves(gs, cs, B, T, H)
            print("    saves: gates %.2e c %.2e" % (rel(g1, gates0), rel(c1, cseq0)), flush=True)
        dy = torch.randn(B, T, H, device=dev)
        dhT, dcT = torch.randn(B, H, device=dev), torch.randn(B, H, device=dev)
        # old BPTT from the OLD saves, new BPTT from the NEW saves (bf16 gates)
        dg0, dh00, dc00 = ops.lstm_tc_bwd(dy, gates0, cseq0, c0, whhT16, dhT, dcT)
        dg1, dh01, dc01 = ops.lstm_c4_bwd(dy, gs, cs, c0, whhT16, dhT, dcT)
        torch.cuda.synchronize()
        print("bwd B%d T%d H%d: dg %.2e dh0 %.2e dc0 %.2e" % (B, T, H, rel(dg1.float(), dg0.float()), rel(dh01, dh00),
                                                               rel(dc01, dc00)), flush=True)

for mode in ([0] if "time" in what else []):
    B, T, H = 32, 500, 1024
    torch.manual_seed(0)

    def mk():
        xg = torch.randn(B, T, 4 * H, device=dev)
        w = (torch.randn(4 * H, H, device=dev) / 32).bfloat16()
        return xg, w, w.t().contiguous()
    sets = [mk(), mk()]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    saved = []
    for xg, w, wT in sets:
        y, hp, hT, cT, gs, cs = ops.lstm_c4_fwd(xg, w, None, None, True)
        saved.append((torch.randn_like(y), gs, cs, wT))
    torch.cuda.synchronize()

    def run(kind, n):
        main = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            s = streams[i]
            s.wait_stream(main)
            with torch.cuda.stream(s):
                for _ in range(3):
                    if kind == "fwd":
                        ops.lstm_c4_fwd(sets[i][0], sets[i][1], None, None, True)
                    else:
                        dy, gs, cs, wT = saved[i]
                        ops.lstm_c4_bwd(dy, gs, cs, None, wT, None, None)
        for i in range(n):
            main.wait_stream(streams[i])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 3 / T * 1e3
    for kind in ("fwd", "bwd"):
        for n in (1, 2):
            run(kind, n)
            print("c4", kind, "streams", n, "%.2f us per step (wall, per layer-step of ONE stream)" % run(kind, n), flush=True)

if "occ" in what:
    for H in (256, 512, 1024):
        print("H", H, "max clusters fwd/4:", lib().eb_lstm_c4_max_clusters(H, 0), "bwd/4:", lib().eb_lstm_c4_max_clusters(H, 4),
              "bwd/8:", lib().eb_lstm_c4_max_clusters(H, 8), "need", H // 32, H // 32, H // 64)
for mode in ([0] if "trace" in what else []):
    B, T, H = 32, 400, 1024
    torch.manual_seed(0)
    xg = torch.randn(B, T, 4 * H, device=dev)
    w = (torch.randn(4 * H, H, device=dev) / 32).bfloat16()
    y_ref = ops.lstm_tc_fwd(xg, w, None, None, False)[0]
    print("  y vs mma.sync kernel: %.2e" % rel(ops.lstm_c4_fwd(xg, w, None, None, True)[0], y_ref))
    tr = torch.zeros(T, 16, dtype=torch.int64, device=dev)
    lib().eb_lstm_c4_set_trace(tr.data_ptr(), T)
    ops.lstm_c4_fwd(xg, w, None, None, True)
    torch.cuda.synchronize()
    lib().eb_lstm_c4_set_trace(None, 0)
    t = tr.cpu().double()[50:350]
    names = {0: "grid barrier passed (block)", 1: "h slice in smem + sync", 2: "mma issued + commit", 3: "acc ready",
             4: "tmem ld, tiles staged, bulk copies issued", 5: "rbar passed (3 remote tiles)", 6: "gates + h store",
             7: "block sync", 8: "threadfence", 9: "(pull loads/stores done)"}
    base = t[:, 0]
    for i, n in names.items():
        print("  +%7.0f cyc  %s" % (float((t[:, i] - base).mean()), n))
    per = float((t[1:, 0] - t[:-1, 0]).mean())
    print("  step period %.0f cyc = %.2f us at 1.965 GHz" % (per, per / 1965))
    print("  threadfence done -> next barrier passed: %.0f cyc" % float((t[1:, 0] - t[:-1, 8]).mean()))


if "trace_bwd" in what:
    B, T, H = 32, 400, 1024
    torch.manual_seed(0)
    xg = torch.randn(B, T, 4 * H, device=dev)
    w = (torch.randn(4 * H, H, device=dev) / 32).bfloat16()
    wT = w.t().contiguous()
    y0, y16, hT0, cT0, gates0, cseq0 = ops.lstm_tc_fwd(xg, w, None, None, True)
    dy = torch.randn_like(y0)
    ops.lstm_tc_bwd(dy, gates0, cseq0, None, wT, None, None)
    tr = torch.zeros(T, 16, dtype=torch.int64, device=dev)
    lib().eb_lstm_tc_set_trace(tr.data_ptr(), T)
    ops.lstm_tc_bwd(dy, gates0, cseq0, None, wT, None, None)
    torch.cuda.synchronize()
    lib().eb_lstm_tc_set_trace(None, 0)
    t = tr.cpu().double()[50:350]
    names = {0: "step top", 1: "phase A + exchange store + sync", 2: "threadfence + atomic", 3: "poll ok (prefetch issued before)", 4: "block sync",
             5: "pull done (warp 0)", 6: "hmma done", 7: "red stored + sync", 8: "reduce -> part", 9: "cluster.sync", 10: "dsmem read (dh)"}
    base = t[:, 0]
    for i, n in names.items():
        print("  +%7.0f cyc  %s" % (float((t[:, i] - base).mean()), n))
    per = float((t[1:, 0] - t[:-1, 0]).mean())
    print("  old BPTT step period %.0f cyc = %.2f us at 1.965 GHz" % (per, per / 1965))


if "skew" in what:
    B, T, H = 32, 300, 1024
    torch.manual_seed(0)
    xg = torch.randn(B, T, 4 * H, device=dev)
    w = (torch.randn(4 * H, H, device=dev) / 32).bfloat16()
    ops.lstm_c4_fwd(xg, w, None, None, True)
    NS = 200
    tr = torch.zeros(NS, 128, 16, dtype=torch.int64, device=dev)
    lib().eb_lstm_c4_set_trace(tr.data_ptr(), -NS)
    ops.lstm_c4_fwd(xg, w, None, None, True)
    torch.cuda.synchronize()
    lib().eb_lstm_c4_set_trace(None, 0)
    t = tr.cpu().double()[50:NS]                       # [steps, cta, slot]  (ns)
    # per step: when did each CTA pass the barrier (0), finish the pull (1), get acc (3), pass rbar (5), store h (6), fence done (8)
    t0 = t[:, :, 0].min(dim=1, keepdim=True).values
    def col(i):
        return (t[:, :, i] - t0)
    import numpy as np
    names = {0: "barrier passed", 1: "pull + sync", 3: "acc ready", 4: "tiles staged", 5: "rbar passed", 6: "h stored", 8: "fence done"}
    print("per-CTA time since the FIRST CTA passed the barrier of the step (ns): mean over steps, then min / median / max over CTAs")
    for i, n in names.items():
        m = col(i).mean(0).numpy()
        print("  %-16s min %6.0f  med %6.0f  max %6.0f   slowest CTAs %s" % (n, m.min(), np.median(m), m.max(), np.argsort(-m)[:6].tolist()))
    per = (t[1:, :, 0] - t[:-1, :, 0]).mean().item()
    print("  step period %.0f ns" % per)
    dur = {"pull": (1, 0), "mma": (3, 1), "stage": (4, 3), "dsmem": (5, 4), "gates": (6, 5), "fence": (8, 6)}
    for n, (a, b) in dur.items():
        m = (t[:, :, a] - t[:, :, b]).mean(0).numpy()
        print("  phase %-6s per CTA: min %6.0f med %6.0f max %6.0f ns" % (n, m.min(), np.median(m), m.max()))
    late = col(8).mean(0).numpy()
    print("  fence-done lateness by cluster rank:", [round(float(late[r::4].mean())) for r in range(4)])
    print("  fence-done lateness, CTAs 0..127 in 8 groups of 16:", [round(float(late[i*16:(i+1)*16].mean())) for i in range(8)])
