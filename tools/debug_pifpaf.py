import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from hyperpose_b200 import capi, models, synthetic as syn
g = models.resnet50_pifpaf(0)
H = W = 129; N = 2
frames = syn.make_frames_u8(8, N, H, W)
eng = capi.Engine(g.to_pack(), (W, H), max_batch_size=N)
eng.infer_u8(frames)
pif, paf = eng.read_outputs(N)
pif = pif.reshape(N, 17, 5, 17, 17); paf = paf.reshape(N, 19, 9, 17, 17)
print('pif conf range', pif[:, :, 0].min(), pif[:, :, 0].max(), 'seeds>0.3', (pif[:, :, 0] > 0.3).sum(), 'paf conf>0.2', (paf[:, :, 0] > 0.2).sum())
dec = capi.PifPafParser(H, W, 0.1)
got = dec.process_batch(pif, paf)
for i in range(N):
    want = oracle.ref_pifpaf_process(pif[i], paf[i], H, W, 0.1)
    print('frame', i, 'gpu', len(got[i]), 'ref', len(want), dec.debug_counts(i))
    for j in range(min(len(got[i]), len(want), 6)):
        a, b = got[i][j], want[j]
        same = a.tobytes() == b.tobytes()
        print('  ', j, same, 'score', a['score'], b['score'], 'nparts', a['parts']['has_value'].sum(), b['parts']['has_value'].sum())
        if not same:
            for k in range(18):
                if a['parts'][k].tobytes() != b['parts'][k].tobytes(): print('      part', k, a['parts'][k], b['parts'][k])

def hr_numpy(pf, h, w):
    HR, WR = (h-1)*8+1, (w-1)*8+1
    m = np.zeros((HR, WR), np.float32)
    f32 = np.float32
    for j in range(h*w):
        c = pf[0].reshape(-1)[j]
        if not c > 0.1: continue
        cx = f32(pf[1].reshape(-1)[j] * f32(8)); cy = f32(pf[2].reshape(-1)[j] * f32(8))
        cs = f32(max(1.0, 0.5 * float(pf[4].reshape(-1)[j]) * 8.0)); cv = f32(c * f32(0.0625)); tc = cs
        clip = lambda v, lo, hi: max(f32(lo), min(f32(hi), f32(v)))
        minx = int(clip(cx - tc, 0, WR-1)); maxx = int(clip(f32(f32(cx + tc) + f32(1)), minx+1, WR))
        miny = int(clip(cy - tc, 0, HR-1)); maxy = int(clip(f32(f32(cy + tc) + f32(1)), miny+1, HR))
        for xx in range(minx, maxx):
            dx2 = f32(f32(xx) - cx); dx2 = f32(dx2*dx2)
            for yy in range(miny, maxy):
                dy2 = f32(f32(yy) - cy); dy2 = f32(dy2*dy2)
                if f32(dx2+dy2) > f32(tc*tc): continue
                if dx2 < 0.25 and dy2 < 0.25: vv = cv
                else:
                    x = f32(-0.5 * float(f32(dx2+dy2)) / float(f32(cs*cs)))
                    if x > 2 or x < -2: e = f32(0)
                    else:
                        e = f32(f32(1) + f32(x / f32(8))); e = f32(e*e); e = f32(e*e); e = f32(e*e)
                    vv = f32(cv * e)
                m[yy, xx] = min(f32(1.0), f32(m[yy, xx] + vv))
    return m
for fld in (0, 5):
    g_hr = dec.debug_hr(0, fld, 17, 17)
    n_hr = hr_numpy(pif[0, fld], 17, 17)
    print('field', fld, 'gpu hr max', g_hr.max(), 'sum', g_hr.sum(), 'numpy max', n_hr.max(), 'sum', n_hr.sum(), 'equal', np.array_equal(g_hr, n_hr), 'nbad', (g_hr != n_hr).sum())
