#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported read-only through oracle/ref_shims.py) with the seeded
synthetic state_dict of oracle/synth.py.  Run in the BUILD container only:

    python tests/golden/make_golden.py

Outputs (committed):
    lhn_256.npz     configs 1-2 + notebook KAT on test_imgs/mortar_pestle.jpg @256
    lhn_dist_256.npz  ColorizeImageTorchDist sample (config 5 semantics)
    lhn_64.npz      small synthetic case incl. per-layer checksums (fast CPU check)
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import synth, ref_shims  # noqa: E402

SEED = 1234


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    CI = ref_shims.import_reference_wrapper()
    model = ref_shims.import_reference_model()
    tmp = tempfile.mkdtemp()
    wpath = os.path.join(tmp, "synthetic_%d.pth" % SEED)
    torch.save(synth.torch_state_dict(SEED), wpath)
    img_path = os.path.join(ref_shims.REF_ROOT, "test_imgs", "mortar_pestle.jpg")

    # ---------------- 256x256, ColorizeImageTorch (configs 1, 2, notebook KAT) -------------
    out = {}
    for maskcent in (False, True):
        tag = "mc1" if maskcent else "mc0"
        cm = CI.ColorizeImageTorch(Xd=256, maskcent=maskcent)
        cm.prep_net(path=wpath)
        cm.load_image(img_path)
        if not maskcent:
            out["img_rgb"] = cm.img_rgb.copy()                        # uint8 256x256x3 (cv2 resize)
            out["img_l_mc"] = cm.img_l_mc.astype(np.float64)
        cases = {}
        cases["zero"] = (np.zeros((2, 256, 256)), np.zeros((1, 256, 256)))
        cases["rand5"] = synth.synthetic_hints(256, 5, 0)
        ab, m = np.zeros((2, 256, 256)), np.zeros((1, 256, 256))
        synth.put_point(ab, m, [135, 160], 3, [23, -69])
        synth.put_point(ab, m, [100, 160], 3, [0, 0])
        cases["kat"] = (ab, m)
        for name, (ab, m) in cases.items():
            rgb = cm.net_forward(ab.copy(), m.copy())
            # raw net output: re-run the net exactly as data/colorize_image.py:263 does
            raw = cm.net.forward(cm.img_l_mc, cm.input_ab_mc, cm.input_mask_mult, cm.mask_cent)[0].cpu().data.numpy()
            if maskcent and name != "rand5":
                continue                                              # keep the fixture small
            out["%s_%s_ab_raw" % (tag, name)] = raw.astype(np.float32)
            if not maskcent:
                out["%s_%s_rgb" % (tag, name)] = rgb.copy()
            if tag == "mc0" and name == "kat":
                out["%s_%s_output_ab" % (tag, name)] = cm.output_ab.astype(np.float32)
            if tag == "mc0" and name == "kat":
                out["kat_fullres_rgb_small"] = cm.get_img_fullres()[::8, ::8].copy()
    np.savez_compressed(os.path.join(HERE, "lhn_256.npz"), **out)
    print("lhn_256.npz", {k: v.shape for k, v in out.items()})

    # ---------------- dist model ------------------------------------------------------------
    cd = CI.ColorizeImageTorchDist(Xd=256, maskcent=True)
    cd.prep_net(path=wpath, dist=True)
    cd.load_image(img_path)
    ab, m = synth.synthetic_hints(256, 5, 0)
    ret = cd.net_forward(ab.copy(), m.copy())
    d = cd.dist_ab                                                     # [529,256,256] float32
    dd = {"ret_quirk": ret.astype(np.float32),
          "dist_rows": d[:, ::4, ::4][:, ::8, ::8].astype(np.float32),   # [529,8,8] of the 64x64 grid
          "dist_upsample_ok": np.array(np.all(d == np.repeat(np.repeat(d[:, ::4, ::4], 4, 1), 4, 2))),
          "dist_sum64": d[:, ::4, ::4].sum(0).astype(np.float32),
          "dist_argmax64": d[:, ::4, ::4].argmax(0).astype(np.int32),
          "dist_max64": d[:, ::4, ::4].max(0).astype(np.float32)}
    np.random.seed(0)
    dd["reccs_128_128_K9"] = cd.get_ab_reccs(128, 128, K=9, N=25000)
    np.savez_compressed(os.path.join(HERE, "lhn_dist_256.npz"), **dd)
    print("lhn_dist_256.npz", {k: v.shape for k, v in dd.items()})

    # ---------------- 64x64 synthetic with per-layer statistics ------------------------------
    L, ab, m = synth.synthetic_batch(2, 64, seed=100, max_hints=4)
    net = model.SIGGRAPHGenerator(dist=True)
    net.load_state_dict(torch.load(wpath))
    net.eval()
    inter = {}
    names = ["model1", "model2", "model3", "model4", "model5", "model6", "model7", "model8", "model9", "model10"]
    hooks = [getattr(net, n).register_forward_hook(lambda mod, i, o, n=n: inter.__setitem__(n, o.detach().numpy().copy()))
             for n in names]
    small = {"L": L, "ab": ab, "mask": m}
    for i in range(2):
        reg, dist = net.forward(L[i], ab[i], m[i], 0.5)
        small["reg_quirk_%d" % i] = reg[0].detach().numpy().astype(np.float32)
        small["dist16_%d" % i] = dist[0, :, ::4, ::4].detach().numpy().astype(np.float32)
        for n in names:
            t = inter[n][0].astype(np.float32)
            small["%s_%d_c8" % (n, i)] = t[:8].copy()                  # first 8 channels, full plane
            small["%s_%d_chmean" % (n, i)] = t.mean(axis=(1, 2))       # every channel, spatial mean
            small["%s_%d_absmax" % (n, i)] = np.abs(t).max(axis=(1, 2))
    for h in hooks:
        h.remove()
    np.savez_compressed(os.path.join(HERE, "lhn_64.npz"), **small)
    print("lhn_64.npz", {k: v.shape for k, v in small.items()})


if __name__ == "__main__":
    main()
