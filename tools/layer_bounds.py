"""Per-layer bound model of the cfg3 conv stack against the measured per-op times (profiles/r01_s2_layers_cfg3.json).

For every conv launch: which kernel the engine picks (mirrors engine.cu's plan logic), the work per k-step, and four lower
bounds per work unit in SM cycles --
  tensor : UMMA floor max(M,128) * N / 256 cycles per K=16 slice (B300_MICROARCH.md, tcgen05 floor), x 4 slices per k-step
  issue  : scalar/uniform instructions of the one MMA-issuing thread per k-step x ~10 cycles (ncu: 44 per k-step in the
           im2col-mode loops, 22 per tap in the unrolled halo loop)
  L2     : bytes the L2 slices must put out per k-step for this SM -- the A / pixel tile in full, the weight tile divided by ~4.5
           (all SMs ask for the same weight tiles at about the same time; ncu: xbar2l1tex 15.5 TB/s for lts2xbar 11.3 TB/s) --
           at 48 B/clk per SM, the chip-wide LTS output cap (~6300 B/clk ~ 11.3-11.9 TB/s) shared by 148 SMs
  epilogue: ~570 cycles per 16 accumulator columns x 128 rows for the 4 epilogue warps (2285 cycles per 64-column tile measured
           on conv1_2 with ncu)
-- and the measured cycles per unit at the run's median SM clock.  Prints a markdown table.
usage: python tools/layer_bounds.py [layers.json] [sm_mhz]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperpose_b200 import models  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r01_s2_layers_cfg3.json")
mhz = float(sys.argv[2]) if len(sys.argv) > 2 else 1657.5
layers = {l["name"]: l for l in json.load(open(path))["layers"]}
g = models.openpose_vgg19(0)
H0, W0, B, SMS = 368, 656, 16, 148


def pick_bn(c):
    for t in (16, 32, 48, 64, 96, 128):
        if c <= t:
            return t
    return 256 if c % 256 == 0 else (128 if c % 128 == 0 else 256)


rows = []
for op in g.ops:
    if op.type != models.OP_CONV:
        continue
    down = g.buffers[op.in_buf][1]
    H, W = H0 >> down, W0 >> down
    px = B * H * W
    R, S, G = op.R, op.S, op.groups
    im2col = op.im2col_input != 0
    ecin = 64 if im2col else (op.cin_g + 63) // 64 * 64
    eR, eS = (1, 1) if im2col else (R, S)
    BN = (op.cout_g + 63) // 64 * 64 if (im2col and op.cout_g <= 128) else pick_bn(op.cout_g)
    cpad = (op.cout_g + BN - 1) // BN * BN
    ksteps = eR * eS * (ecin // 64)
    tma_store = op.out_mode == models.OUT_F16_NHWC and BN % 64 == 0 and (G == 1 or op.cout_g % 64 == 0)
    swap = tma_store and cpad == 128 and op.cout_g == 128 and ksteps >= 18
    ty, tx = (H + 15) // 16, (W + 7) // 8
    waste = ty * 16 * tx * 8 / (H * W) - 1
    halo = (not im2col) and eR == 3 and eS == 3 and tma_store and waste <= 0.06 and not swap
    ms = layers[op.name]["ms"]
    if im2col:
        kern, units, M, N = "stem3", (px + 127) // 128, 128, BN
        tensor, issue, ingest, epi = 2 * (128 * N / 256), 60, 0, (N / 16) * 570
        ksteps_u = 1
    elif swap:
        npx = 208 if (H, W) == (46, 82) else 256
        kern, units = f"swap N={npx}", ((px + npx - 1) // npx) * G
        tensor, issue, ingest = 4 * (128 * npx / 256), 440, (16384 / 4.5 + npx * 128) / 48
        epi, ksteps_u = (npx / 16) * 300, ksteps
    elif halo:
        resident = G == 1 and cpad == BN and (1024 + 2 * 23552 + ksteps * BN * 128 + 32768 + 400) <= 227 * 1024
        kern = "halo" + ("+resident W" if resident else "")
        units = B * ty * tx * G * (cpad // BN)
        tensor, issue = 4 * (128 * BN / 256), 220
        ingest = ((23040 / 9) + (0 if resident else BN * 128 / 4.5)) / 48
        epi, ksteps_u = (BN / 16) * 570, ksteps
    else:
        kern, units = f"im2col BN={BN}", ((px + 127) // 128) * G * (cpad // BN)
        tensor, issue, ingest = 4 * (128 * BN / 256), 440, (16384 + BN * 128 / 4.5) / 48
        epi, ksteps_u = (BN / 16) * 570, ksteps
    waves = units / SMS
    rounds = -(-units // SMS)
    meas = ms * 1e-3 * mhz * 1e6 / rounds            # cycles per unit on the critical SM
    main = max(tensor, issue, ingest) * ksteps_u
    bound = max(("tensor", tensor * ksteps_u), ("issue", issue * ksteps_u), ("L2", ingest * ksteps_u), ("epilogue", epi), key=lambda t: t[1])
    rows.append((op.name, kern, f"{H}x{W}", ksteps_u, f"{waves:.2f}", ms, layers[op.name]["tflops"], tensor * ksteps_u, issue * ksteps_u, ingest * ksteps_u, epi, meas, bound[0], max(main, epi) / meas))

print("| layer | kernel | map | k-steps | waves | ms | TF/s | tensor | issue | L2 | epilogue | measured | largest bound | bound / measured |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]} | {r[4]} | {r[5]:.4f} | {r[6]:.0f} | {r[7]:.0f} | {r[8]:.0f} | {r[9]:.0f} | {r[10]:.0f} | {r[11]:.0f} | {r[12]} | {r[13]:.2f} |")
tot = sum(r[5] for r in rows)
print(f"\nconv launches: {len(rows)}, {tot:.3f} ms per step at {mhz:.0f} MHz; cycles are per work unit (tile / pixel unit) on one SM, measured = time x clock / rounds of the persistent grid")
