"""Where does a frame's result depend on the other frames of its batch?  Runs the eager units of the model on n frames
and on the two halves and reports the first tensor that is not bit-identical (debug aid for SURVEY.md 8e)."""
import sys

import torch

sys.path.insert(0, ".")
from diffusionvid_b200 import model as pm, ops, synth  # noqa: E402
from oracle import model as om  # noqa: E402

H, W = 192, 256
HP = dict(num_proposals=100, num_classes=30, hidden=256, nheads=8, dim_dynamic=64, dim_ff=2048, num_heads=3,
          num_heads_local=1, num_cls=1, num_reg=3, sample_step=4, snr_scale=2.0, use_nms=True, infer_batch=8,
          all_frame_interval=8, key_frame_location=0, global_enable=True, mem_size=300, mem_size2=50,
          topk=(75, 25), pixel_mean=(123.675, 116.280, 103.530), pixel_std=(58.395, 57.120, 57.375),
          blocks=(2, 2, 3, 2), device="cuda")


def main():
    m = pm.DiffusionDet(HP)
    m.load_state_dict(synth.make_state_dict(seed=21, blocks=HP["blocks"]), strict=False)
    m.to("cuda")
    m.use_graphs = False
    m._pack()
    m._warm_constants([999, 749, 499, 249])
    frames = synth.make_clip(19, H, W, seed=6).cuda()
    noise = om.NoiseSource(9, 100)
    for n, lo in ((8, 8), (13, 0), (8, 0)):
        imgs = frames[lo:lo + n].contiguous()
        binit = noise.get("init", 0, 8, 0, n).cuda() if n <= 8 else torch.cat([noise.get("init", 0, 0, 0, 8), noise.get("init", 0, 0, 1, n - 8)]).cuda()
        full = m._extract(imgs, binit.contiguous(), W, H)
        torch.cuda.synchronize()
        for parts in ([(0, n // 2), (n // 2, n)], [(i, i + 1) for i in range(n)]):
            bad = {}
            for a, b in parts:
                sub = m._extract(imgs[a:b].contiguous(), binit[a:b].contiguous(), W, H)
                torch.cuda.synchronize()
                for k in ("p3", "p4", "p5", "lg", "bx", "o32", "o16", "k1", "k2"):
                    if not torch.equal(sub[k], full[k][a:b]):
                        d = (sub[k].float() - full[k][a:b].float()).abs().max().item()
                        bad.setdefault(k, []).append(((a, b), d))
            print("n=%d lo=%d parts=%s ->" % (n, lo, parts[:2]), {k: v[:3] for k, v in bad.items()} or "bit-identical")
    # per-layer walk of the backbone for the failing case
    imgs = frames[8:16].contiguous()
    real = ops.conv2d
    log = []

    def spy(x, w, b, cout, R, S, st, pad, relu, resid=None, resid_shift=0, out=None):
        y = real(x, w, b, cout, R, S, st, pad, relu, resid=resid, resid_shift=resid_shift, out=out)
        log.append((tuple(x.shape), cout, R, st, y.clone()))
        return y
    ops.conv2d = spy
    log.clear(); m.extract_features(imgs); full_log = list(log)
    log.clear(); m.extract_features(imgs[:4].contiguous()); half_log = list(log)
    ops.conv2d = real
    for i, (fl, hl) in enumerate(zip(full_log, half_log)):
        if not torch.equal(fl[4][:4], hl[4]):
            print("first differing conv: #%d in%s cout=%d R=%d stride=%d maxdiff=%g" %
                  (i, fl[0], fl[1], fl[2], fl[3], (fl[4][:4].float() - hl[4].float()).abs().max().item()))
            break
    else:
        print("all %d convs bit-identical between 8 frames and the first 4" % len(full_log))


if __name__ == "__main__":
    with torch.no_grad():
        main()
