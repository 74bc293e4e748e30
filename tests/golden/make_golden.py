#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference code.

Run in the build container only (needs /root/reference, read-only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

* `function.py` and `net.py` are imported from
  /root/reference/style_transfer/AdaIN unchanged.
* `style_transfer` / `calc_sum` are FunctionDef nodes AST-extracted from
  CCST_OverallStyleTransfer.py / mean_std_computation_effcientMem.py /
  CCST_SingleStyleTransfer.py (the scripts cannot be imported: argparse and
  torch.load run at module level), compiled and exec'd as they are.
* the accumulate / finalise statements of the scripts
  (mean_std_computation_effcientMem.py:129-131,135-137 and
  CCST_SingleStyleTransfer.py:201-203) are exec'd from their source lines.

Inputs and weights come from ccst_b200.synth (seeded, CPU generator).  The
resulting vectors travel with the repo; the reference does not.
"""
from __future__ import annotations

import ast
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CCST_REF", "/root/reference/style_transfer/AdaIN")
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import function as ref_function  # noqa: E402  (reference)
import net as ref_net  # noqa: E402  (reference)

from ccst_b200 import synth  # noqa: E402


def _extract_funcs(path, names, ns):
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), ns)
    return ns


def _source_lines(path, first, last):
    import textwrap
    lines = open(path).read().splitlines()[first - 1:last]
    return textwrap.dedent("\n".join(lines)) + "\n"


def ref_models(seed=0):
    """Reference nn.Sequential objects carrying the synthetic weights."""
    vgg_full = ref_net.vgg
    dec = ref_net.decoder
    synth.init_vgg_(vgg_full, seed)
    synth.init_decoder_(dec, seed)
    vgg = torch.nn.Sequential(*list(vgg_full.children())[:31])
    return vgg.eval(), dec.eval()


def weights_digest(*mods):
    h = hashlib.sha256()
    for m in mods:
        for k, v in m.state_dict().items():
            h.update(k.encode())
            h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def main():
    torch.set_num_threads(1)  # deterministic reduction order
    out = {}
    overall_py = os.path.join(REF, "CCST_OverallStyleTransfer.py")
    single_py = os.path.join(REF, "CCST_SingleStyleTransfer.py")
    stats_py = os.path.join(REF, "mean_std_computation_effcientMem.py")

    ns_overall = {"torch": torch, "device": torch.device("cpu"),
                  "adaIN_StyleStat_ContentFeat": ref_function.adaIN_StyleStat_ContentFeat}
    _extract_funcs(overall_py, {"style_transfer"}, ns_overall)
    ns_single = dict(ns_overall)
    _extract_funcs(single_py, {"style_transfer", "calc_sum"}, ns_single)
    ns_stats = {"torch": torch}
    _extract_funcs(stats_py, {"calc_sum", "calc_mean_std"}, ns_stats)

    # ---------------- statistics ----------------
    stats = {}
    cases = {
        "s_small": synth.features((3, 16, 9, 7), 11),
        "s_12x12": synth.features((4, 32, 12, 12), 12),
        "s_64x64": synth.features((2, 8, 64, 64), 13),
        "s_odd": synth.features((2, 5, 7, 7), 14),
        "s_bigmean": synth.features((2, 6, 16, 16), 15, relu=False) + 300.0,
        "s_hw1": synth.features((2, 4, 1, 1), 16),
    }
    const = synth.features((1, 4, 8, 8), 17)
    const[:, 1] = 0.0
    const[:, 2] = 2.5
    cases["s_const"] = const
    for name, feat in cases.items():
        m, s = ref_function.calc_mean_std(feat)
        stats[name + "/x"] = feat.numpy()
        stats[name + "/mean"] = m.numpy()
        stats[name + "/std"] = s.numpy()
        m64, s64 = ref_function.calc_mean_std(feat.double())
        stats[name + "/mean64"] = m64.numpy()
        stats[name + "/std64"] = s64.numpy()
    np.savez_compressed(os.path.join(HERE, "stats.npz"), **stats)

    # ---------------- AdaIN ----------------
    adain = {}
    c = synth.features((3, 16, 10, 12), 21)
    s = synth.features((3, 16, 7, 9), 22)
    adain["feat/content"] = c.numpy()
    adain["feat/style"] = s.numpy()
    adain["feat/out"] = ref_function.adaptive_instance_normalization(c, s).numpy()
    sm = torch.randn((1, 16, 1, 1), generator=torch.Generator().manual_seed(23))
    ss = torch.rand((1, 16, 1, 1), generator=torch.Generator().manual_seed(24)) + 0.1
    adain["stat/content"] = c.numpy()
    adain["stat/mean"] = sm.numpy()
    adain["stat/std"] = ss.numpy()
    adain["stat/out"] = ref_function.adaIN_StyleStat_ContentFeat(c, [sm, ss]).numpy()
    c2 = synth.features((2, 8, 64, 64), 25)
    sm2 = torch.randn((1, 8, 1, 1), generator=torch.Generator().manual_seed(26))
    ss2 = torch.rand((1, 8, 1, 1), generator=torch.Generator().manual_seed(27)) + 0.1
    adain["stat64/content"] = c2.numpy()
    adain["stat64/mean"] = sm2.numpy()
    adain["stat64/std"] = ss2.numpy()
    adain["stat64/out"] = ref_function.adaIN_StyleStat_ContentFeat(c2, [sm2, ss2]).numpy()
    np.savez_compressed(os.path.join(HERE, "adain.npz"), **adain)

    # ---------------- calc_sum + accumulation + finalise ----------------
    acc = {}
    batches = [synth.features((3, 12, 8, 8), 31), synth.features((3, 12, 8, 8), 32),
               synth.features((2, 12, 8, 8), 33)]
    accumulate_src = _source_lines(stats_py, 129, 131)
    finalize_src = _source_lines(stats_py, 135, 137)
    assert "all_feat_sum += feat_sum" in accumulate_src, accumulate_src
    assert "feat_std = torch.sqrt(feat_var + 1e-5)" in finalize_src, finalize_src
    for tag, cast in (("f32", lambda t: t), ("f64", lambda t: t.double())):
        env = {"torch": torch, "all_feat_sum": 0, "all_feat_square_sum": 0, "all_count": 0}
        for i, b in enumerate(batches):
            fs, fss, cnt = ns_stats["calc_sum"](cast(b))
            if tag == "f32":
                acc[f"b{i}/x"] = b.numpy()
                acc[f"b{i}/sum"] = fs.numpy()
                acc[f"b{i}/sqsum"] = fss.numpy()
                acc[f"b{i}/count"] = np.int64(cnt)
            env.update(feat_sum=fs, feat_square_sum=fss, count=cnt)
            exec(accumulate_src, env)
        exec(finalize_src, env)
        acc[f"final_{tag}/mean"] = env["feat_mean"].numpy()
        acc[f"final_{tag}/std"] = env["feat_std"].numpy()
        acc[f"final_{tag}/count"] = np.int64(env["all_count"])
    # the np.save payload of :146
    acc["npy_payload"] = np.asarray([acc["final_f32/mean"], acc["final_f32/std"]])
    np.savez_compressed(os.path.join(HERE, "overall_stats.npz"), **acc)

    # ---------------- encoder / decoder / style_transfer ----------------
    vgg, dec = ref_models(0)
    net = {"weights_sha256": np.frombuffer(weights_digest(vgg, dec).encode(), dtype=np.uint8)}
    single_final_src = _source_lines(single_py, 201, 203)
    assert "feat_std = torch.sqrt(feat_var + 1e-5)" in single_final_src
    with torch.no_grad():
        for tag, (n, h, w) in {"sq40": (2, 40, 40), "odd37x45": (1, 37, 45), "r96": (2, 96, 96)}.items():
            x = synth.images(n, h, w, 41 + h)
            f = vgg(x)
            style_img = synth.images(1, h + 8, w + 3, 77 + h)
            sf = vgg(style_img)
            env = {"torch": torch}
            fs, fss, cnt = ns_single["calc_sum"](sf)
            env.update(feat_sum=fs, feat_square_sum=fss, count=cnt,
                       feat_mean=fs / float(cnt))
            exec(single_final_src, env)
            style_stat = [env["feat_mean"], env["feat_std"]]
            net[f"{tag}/x"] = x.numpy()
            net[f"{tag}/style_img"] = style_img.numpy()
            net[f"{tag}/feat"] = f.numpy()
            net[f"{tag}/style_mean"] = style_stat[0].numpy()
            net[f"{tag}/style_std"] = style_stat[1].numpy()
            net[f"{tag}/dec_of_feat"] = dec(f).numpy()
            for alpha in (1.0, 0.6):
                o = ns_overall["style_transfer"](vgg, dec, x, style_stat, alpha)
                net[f"{tag}/out_a{alpha}"] = o.numpy()
                o2 = ns_single["style_transfer"](vgg, dec, x, style_stat, alpha)
                assert torch.equal(o, o2)
            if n >= 2:
                o = ns_overall["style_transfer"](vgg, dec, x, style_stat, 1.0, [0.25, 0.75])
                net[f"{tag}/out_interp"] = o.numpy()
    np.savez_compressed(os.path.join(HERE, "net.npz"), **net)
    for fn in ("stats", "adain", "overall_stats", "net"):
        p = os.path.join(HERE, fn + ".npz")
        print(fn, os.path.getsize(p) // 1024, "KiB")


def make_io():
    """Image I/O either side of style_transfer: the loader transform of
    cjm_util/data_helper.py:45-49 (`Resize((S,S))` + `ToTensor()`) applied by the REAL torchvision to
    PIL images, the reference `style_transfer`, then the REAL `save_image` of
    CCST_OverallStyleTransfer.py:158-167 (PNG in memory, decoded back: lossless) -> io_u8.npz."""
    import io

    from PIL import Image
    from torchvision import transforms
    from torchvision.utils import save_image

    torch.set_num_threads(1)
    overall_py = os.path.join(REF, "CCST_OverallStyleTransfer.py")
    ns = {"torch": torch, "device": torch.device("cpu"),
          "adaIN_StyleStat_ContentFeat": ref_function.adaIN_StyleStat_ContentFeat}
    _extract_funcs(overall_py, {"style_transfer"}, ns)
    vgg, dec = ref_models(0)
    out = {"weights_sha256": np.frombuffer(weights_digest(vgg, dec).encode(), dtype=np.uint8)}
    n, h, w = 2, 40, 48
    g = torch.Generator().manual_seed(4242)
    # smooth-ish uint8 images covering the full 0..255 range
    base = torch.rand((n, h // 4, w // 4, 3), generator=g)
    img = torch.nn.functional.interpolate(base.permute(0, 3, 1, 2), size=(h, w), mode="bilinear",
                                          align_corners=False).permute(0, 2, 3, 1)
    img = (img + 0.15 * torch.rand((n, h, w, 3), generator=g)).clamp(0, 1)
    x_u8 = (img * 255).round().to(torch.uint8).numpy()
    x_u8[0, 0, 0] = (0, 255, 128)
    tr = transforms.Compose([transforms.Resize((h, w)), transforms.ToTensor()])
    x = torch.stack([tr(Image.fromarray(x_u8[i])) for i in range(n)])
    # style statistics of an encoded style image through the reference's own single-style lines
    # (CCST_SingleStyleTransfer.py:55-67 calc_sum, :201-203): relu4_1-like statistics keep the decoder's
    # output inside [0, 1], the range the image tolerance of BASELINE.json is stated for
    single_py = os.path.join(REF, "CCST_SingleStyleTransfer.py")
    ns_single = {"torch": torch}
    _extract_funcs(single_py, {"calc_sum"}, ns_single)
    single_final_src = _source_lines(single_py, 201, 203)
    assert "feat_std = torch.sqrt(feat_var + 1e-5)" in single_final_src
    with torch.no_grad():
        sf = vgg(synth.images(1, h + 8, w - 8, 4243))
    fs, fss, cnt = ns_single["calc_sum"](sf)
    env = {"torch": torch, "feat_sum": fs, "feat_square_sum": fss, "count": cnt, "feat_mean": fs / float(cnt)}
    exec(single_final_src, env)
    sm, ss = env["feat_mean"], env["feat_std"]
    out["x_u8"] = x_u8
    out["x_tensor"] = x.numpy()
    out["style_mean"] = sm.numpy()
    out["style_std"] = ss.numpy()
    with torch.no_grad():
        for alpha in (1.0, 0.5):
            o = ns["style_transfer"](vgg, dec, x, [sm, ss], alpha)
            # stretch so that both clamps of save_image are exercised
            o = (o - 0.5) * 3.0 + 0.5 if alpha == 0.5 else o
            saved = []
            for out_img in o:
                buf = io.BytesIO()
                save_image(out_img, buf, format="png")
                buf.seek(0)
                saved.append(np.asarray(Image.open(buf).convert("RGB")))
            out[f"out_f32_a{alpha}"] = o.numpy()
            out[f"out_u8_a{alpha}"] = np.stack(saved)
        # `resize = transforms.Resize(args.output_size); output = resize(output)` (:134-135,154-155),
        # the real torchvision transform on the reference's output, then save_image
        o = ns["style_transfer"](vgg, dec, x, [sm, ss], 1.0)
        for tag, size in (("s24", 24), ("s17", 17), ("s64", 64), ("hw", (20, 31))):
            r = transforms.Resize(size)(o)
            out[f"resize_{tag}"] = r.numpy()
            saved = []
            for out_img in r:
                buf = io.BytesIO()
                save_image(out_img, buf, format="png")
                buf.seek(0)
                saved.append(np.asarray(Image.open(buf).convert("RGB")))
            out[f"resize_{tag}_u8"] = np.stack(saved)
        # the batch-wide calc_mean_std of mean_std_computation_effcientMem.py:89-101 (AST-extracted, as is)
        ns_stats = {"torch": torch}
        _extract_funcs(os.path.join(REF, "mean_std_computation_effcientMem.py"), {"calc_mean_std"}, ns_stats)
        bf = synth.features((3, 24, 9, 11), 91)
        bm, bs = ns_stats["calc_mean_std"](bf)
        bm64, bs64 = ns_stats["calc_mean_std"](bf.double())
        out["batchstat/x"], out["batchstat/mean"], out["batchstat/std"] = bf.numpy(), bm.numpy(), bs.numpy()
        out["batchstat/mean64"], out["batchstat/std64"] = bm64.numpy(), bs64.numpy()
        # the Camelyon command line: --image_size 512 --output_size 96 (:187-191), on a synthetic image
        # (input regenerated by the tests from the same seed; only the 96x96 result is stored)
        big = torch.rand((1, 3, 512, 512), generator=torch.Generator().manual_seed(77))
        out["resize_512_to_96"] = transforms.Resize(96)(big).numpy()
    # the loader's Resize((S, S)) on PIL images (data_helper.py:45-49): REAL torchvision / Pillow, up- and
    # down-sizing, non-square inputs; smooth-ish inputs so that the archive stays small
    import hashlib as _hl
    for tag, (ih, iw, S) in {"up": (37, 53, 64), "down": (150, 100, 48), "pacs": (227, 227, 512)}.items():
        gg = torch.Generator().manual_seed(600 + ih)
        base_ = torch.rand((1, ih // 3 + 1, iw // 3 + 1, 3), generator=gg)
        im_ = torch.nn.functional.interpolate(base_.permute(0, 3, 1, 2), size=(ih, iw), mode="bilinear",
                                              align_corners=False).permute(0, 2, 3, 1)
        im_ = ((im_ + 0.2 * torch.rand((1, ih, iw, 3), generator=gg)).clamp(0, 1) * 255).round().to(torch.uint8).numpy()[0]
        r_ = np.asarray(transforms.Resize((S, S))(Image.fromarray(im_)))
        out[f"resize_in_{tag}/x"] = im_
        if tag == "pacs":  # 786 KB: keep a digest and one row
            out[f"resize_in_{tag}/sha256"] = np.frombuffer(_hl.sha256(r_.tobytes()).hexdigest().encode(), dtype=np.uint8)
            out[f"resize_in_{tag}/row100"] = r_[100]
        else:
            out[f"resize_in_{tag}/y"] = r_
    np.savez_compressed(os.path.join(HERE, "io_u8.npz"), **out)
    print("io_u8", os.path.getsize(os.path.join(HERE, "io_u8.npz")) // 1024, "KiB",
          "a=1 range [%.3f, %.3f]" % (out["out_f32_a1.0"].min(), out["out_f32_a1.0"].max()),
          "clamped lo/hi:", int((out["out_u8_a0.5"] == 0).sum()), int((out["out_u8_a0.5"] == 255).sum()))


def make_f4():
    """SURVEY 8f rank 4, forward only: the REAL `Net` class (net.py:95-152) and the REAL MixStyle module
    (nets/layers.py) run on seeded inputs -> f4.npz."""
    import importlib.util
    import random

    torch.set_num_threads(1)
    out = {}
    vgg_full = ref_net.vgg
    dec = ref_net.decoder
    synth.init_vgg_(vgg_full, 0)
    synth.init_decoder_(dec, 0)
    net = ref_net.Net(torch.nn.Sequential(*list(vgg_full.children())[:31]), dec).eval()
    out["weights_sha256"] = np.frombuffer(weights_digest(net.enc_1, net.enc_2, net.enc_3, net.enc_4, dec).encode(), dtype=np.uint8)
    content = synth.images(2, 64, 72, 901)  # calc_style_loss asserts equal feature sizes: same image size
    style = synth.images(2, 64, 72, 902)
    with torch.no_grad():
        for alpha in (1.0, 0.7):
            lc, ls = net(content, style, alpha)
            out[f"loss_c_a{alpha}"] = np.float32(lc.item())
            out[f"loss_s_a{alpha}"] = np.float32(ls.item())
        feats = net.encode_with_intermediate(style)
        for i, f in enumerate(feats):
            m, sd = ref_function.calc_mean_std(f)
            out[f"style_level{i}_mean"], out[f"style_level{i}_std"] = m.numpy(), sd.numpy()
    out["content"], out["style"] = content.numpy(), style.numpy()
    # MixStyle: the real module, its two random draws replayed from the same seeds
    spec = importlib.util.spec_from_file_location("ref_layers", os.path.join(os.path.dirname(os.path.dirname(REF)), "nets", "layers.py"))
    layers = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(layers)
    mix = layers.MixStyle(p=1.0, alpha=0.1, eps=1e-6, mix="random")
    mix.train()
    x = synth.features((4, 16, 9, 11), 903)
    torch.manual_seed(77)
    random.seed(77)
    y = mix(x)
    torch.manual_seed(77)
    lmda = mix.beta.sample((4, 1, 1, 1))
    perm = torch.randperm(4)
    out["mix_x"], out["mix_y"], out["mix_lmda"], out["mix_perm"] = x.numpy(), y.numpy(), lmda.numpy(), perm.numpy()
    np.savez_compressed(os.path.join(HERE, "f4.npz"), **out)
    print("f4", os.path.getsize(os.path.join(HERE, "f4.npz")) // 1024, "KiB", {k: float(v) for k, v in out.items() if k.startswith("loss")})


if __name__ == "__main__":
    if "--only-f4" in sys.argv:
        make_f4()
    elif "--only-io" in sys.argv:
        make_io()
    else:
        main()
        make_io()
        make_f4()
