"""Inference twins with the reference's public API and behaviour:

    MMEnsemble(ckpt_dir, device).predict(image_path, tab_dict|None) -> (prob, debug)    ui/gradio_app/infer_mm.py:41-109
    MILEnsemble(ckpt_dir, device, backbone).predict(processed_dir)  -> (prob|None, dbg) ui/gradio_app/infer_mil.py:103-193

Differences that are deliberate: the three TTA flips run as ONE batch-3 forward per fold, each fold's forward
is a captured CUDA graph, there is one device->host read per prediction instead of one per fold (q8); MILEnsemble accepts the
trainer's checkpoint layout ('model' key, hid 128) which the reference twin silently fails to load (q9).
"""
from pathlib import Path

import numpy as np
import torch
from PIL import Image

from . import ops  # noqa: F401
from ._lib import init
from .modules import MMNet, MILNetTwin
from .preproc import normalize_flip, resize_center_crop

TAB_FEATURES = ['depth', 'width', 'enamel_cracks', 'occlusal_load', 'carious_lesion',
                'opposing_type', 'adjacent_teeth', 'age_range', 'cervical_lesion']


@torch.no_grad()
def tta_logit(model, x_img, x_tab):
    """mean logit over {identity, W-flip, H-flip} (train_mm_joint_dualtask.py:326-335) as one batched forward."""
    B = x_img.shape[0]
    x3 = torch.cat([x_img, torch.flip(x_img, dims=[3]), torch.flip(x_img, dims=[2])], 0)
    logit, _ = model(x3, x_tab.repeat(3, 1))
    return logit.view(3, B).mean(0)


class _GraphedForward:
    """Captures `fn(*static_inputs)` once and replays it: batch-1/3 EfficientNet forwards are launch-bound otherwise."""

    def __init__(self, fn, example_inputs, warmup=2):
        self.static_in = [t.clone() for t in example_inputs]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for d, s in zip(self.static_in, inputs):
            d.copy_(s, non_blocking=True)
        self.graph.replay()
        return self.static_out


def replay_concurrently(graphed, inputs, streams):
    """Replay several captured forwards (the folds of an ensemble: independent models, same input) at the same time, one
    stream each, forked from and joined back into the current stream.  A batch-3 EfficientNet forward is ~200 dependent
    kernels of a few microseconds that leave most of the GPU idle, so five of them side by side cost little more than one
    (measured: 6.3 ms one after the other).  -> list of the static outputs (valid on the current stream after the join)."""
    main = torch.cuda.current_stream()
    outs = []
    for g, args, st in zip(graphed, inputs, streams):
        st.wait_stream(main)
        with torch.cuda.stream(st):
            outs.append(g(*args))
    for st in streams[:len(graphed)]:
        main.wait_stream(st)
    return outs


def eval_resize_crop(img, size, swap_channels=False):
    """timm eval transform up to the uint8 image: Resize(floor(S/0.875), bicubic) -> CenterCrop(S) (infer_mm.py:12-17),
    on the device and bit-identical to the PIL path (preproc.resize_center_crop).  img: PIL image or uint8 HWC array
    -> CUDA uint8 [S,S,3]."""
    return resize_center_crop(np.asarray(img), int(np.floor(size / 0.875)), size, "bicubic", swap_channels=swap_channels)


class MMEnsemble:
    def __init__(self, ckpt_dir, device='cuda', graph=True):
        self.ckpt_dir = Path(ckpt_dir)
        self.device = device
        self.ckpts = sorted(self.ckpt_dir.glob("mm_dualtask_fold*.pt"))
        self.models, self.scales = [], []
        self.img_size = None
        self.batch_size = None
        self.use_graph = graph
        self._graphs = {}
        self._streams = []
        self._load()

    @property
    def num_folds(self):
        return len(self.models)

    def _load(self):
        for ck in self.ckpts:
            ckpt = torch.load(ck, map_location='cpu', weights_only=False)
            a = ckpt['args']
            model = MMNet(backbone=a['backbone'], tab_in=len(TAB_FEATURES), tab_hidden=a['tab_hidden'], drop=a['dropout'])
            model.load_state_dict(ckpt['model'], strict=True)
            model = model.to(self.device).eval()
            self.models.append((model, float(ckpt['T'])))
            n = len(TAB_FEATURES)
            mean = np.array(ckpt['scaler_mean']) if ckpt.get('scaler_mean') is not None else np.zeros(n)
            scale = np.array(ckpt['scaler_scale']) if ckpt.get('scaler_scale') is not None else np.ones(n)
            self.scales.append((mean, scale))
            self.img_size = a['img_size']
        if not self.models:
            print("[MMEnsemble] No checkpoints found.")
        else:
            init(self.device)

    def _prep_tab(self, tab_dict, fold=0):
        mean, scale = self.scales[fold]
        x = mean.copy() if tab_dict is None else np.array([float(tab_dict[k]) for k in TAB_FEATURES], dtype=np.float32)
        z = (x - mean) / np.where(scale == 0, 1.0, scale)
        return torch.tensor(z, dtype=torch.float32).unsqueeze(0)

    def _fold_logits(self, f, x3, xt3):
        model = self.models[f][0]
        if not self.use_graph:
            return model(x3, xt3)[0]
        if f not in self._graphs:
            self._graphs[f] = _GraphedForward(lambda a, b: model(a, b)[0], [x3, xt3])
        return self._graphs[f](x3, xt3)

    @torch.no_grad()
    def predict_tensor(self, bgr_u8, tab_dict=None):
        """bgr_u8: CUDA uint8 [S,S,3].  Returns a device tensor of per-fold probabilities (no host sync)."""
        with torch.cuda.device(bgr_u8.device):       # graphs and side streams are made on the ensemble's device, not the current one
            x3 = torch.stack([normalize_flip(bgr_u8, f) for f in (0, 1, 2)], 0)      # TTA: identity, W-flip, H-flip
            xts = [self._prep_tab(tab_dict, fold=f).to(bgr_u8.device, non_blocking=True).repeat(3, 1) for f in range(len(self.models))]
            if self.use_graph and all(f in self._graphs for f in range(len(self.models))):
                # the folds are independent: their captured forwards replay side by side, one stream per fold
                if len(self._streams) < len(self.models):
                    self._streams = [torch.cuda.Stream(device=bgr_u8.device) for _ in self.models]
                logits = replay_concurrently([self._graphs[f] for f in range(len(self.models))],
                                             [(x3, xt) for xt in xts], self._streams)
            else:                                    # first call: capture fold by fold
                logits = [self._fold_logits(f, x3, xts[f]) for f in range(len(self.models))]
            probs = [torch.sigmoid(lg.mean(0, keepdim=True) / T) for lg, (_, T) in zip(logits, self.models)]
            return torch.cat(probs)

    def predict_image(self, rgb_u8, tab_dict=None):
        """Decoded image (PIL image or uint8 RGB [H,W,3] host array) -> per-fold probabilities (numpy): full-size upload,
        eval transform + TTA + all folds on the device, one device->host read."""
        with torch.cuda.device(self.device):
            bgr = eval_resize_crop(rgb_u8, self.img_size, swap_channels=True)
        return self.predict_tensor(bgr, tab_dict).cpu().numpy()                      # the one sync of the prediction

    def predict(self, image_path, tab_dict=None):
        """Return (prob_mm, debug_str)."""
        if not self.models:
            return 0.5, "MM not loaded"
        probs = self.predict_image(Image.open(image_path).convert('RGB'), tab_dict)
        return float(np.mean(probs)), f"fold_probs={np.round(probs, 3)}"


def _remap_state_dict_keys(sd):
    """ui/gradio_app/infer_mil.py:17-34"""
    out = {}
    for k, v in sd.items():
        nk = "enc." + k[len("encoder."):] if k.startswith("encoder.") else k
        for n in "VUw":
            nk = nk.replace(f"mil.attention_{n}.", f"mil.{n}.")
        out[nk] = v
    return out


def _list_images(folder_or_file):
    p = Path(folder_or_file)
    exts = {".png", ".jpg", ".jpeg", ".bmp", ".tif", ".tiff", ".webp"}
    if p.is_file():
        return [p] if p.suffix.lower() in exts else []
    if p.is_dir():
        return sorted(q for q in p.iterdir() if q.suffix.lower() in exts)
    return []


class MILEnsemble:
    def __init__(self, ckpt_dir, device="cuda", backbone="tf_efficientnet_b0_ns"):
        self.ckpt_dir = Path(ckpt_dir)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("teethrt MILEnsemble runs on CUDA (sm_100a) only; there is no CPU fallback")
        self.backbone = backbone
        self.num_folds = 0
        self.models = []
        # transforms.Resize(512) (PIL bilinear, antialiased) -> CenterCrop(480) -> ToTensor, on the device (infer_mil.py:116-119)
        self._tfm = lambda im: resize_center_crop(np.asarray(im), 512, 480, "bilinear").permute(2, 0, 1).float().div(255)
        self._load_folds()

    def predict(self, processed_dir):
        if not self.models:
            return None, "MIL not loaded"
        imgs = _list_images(processed_dir)
        if not imgs:
            return None, f"MIL: no images under {processed_dir}"
        bag = []
        for p in imgs:
            try:
                im = Image.open(p).convert("RGB")
            except Exception:
                continue
            with torch.cuda.device(self.device):
                bag.append(self._tfm(im))
        if not bag:
            return None, f"MIL: failed to load images in {processed_dir}"
        x = torch.stack(bag, dim=0)
        with torch.no_grad(), torch.cuda.device(self.device):
            logits = torch.stack([m(x) for m in self.models]).cpu()                  # one sync for all folds
        logit_mean = float(logits.mean())
        prob = float(torch.sigmoid(torch.tensor(logit_mean)))
        dbg = f"Instances={x.shape[0]} | folds={len(self.models)} | logits={[round(float(l), 4) for l in logits]}"
        return prob, dbg

    def _load_folds(self):
        if not self.ckpt_dir.exists():
            return
        for k in range(10):
            cand = self.ckpt_dir / f"mil_v1_fold{k}.pt"
            if not cand.exists():
                continue
            ckpt = torch.load(cand, map_location="cpu", weights_only=False)
            sd = ckpt.get("model", ckpt.get("state_dict", ckpt)) if isinstance(ckpt, dict) else ckpt
            sd = _remap_state_dict_keys(sd)
            hid = sd["mil.V.weight"].shape[0]
            model = MILNetTwin(self.backbone, pretrained=False, hid_dim=hid)
            model.load_state_dict(sd, strict=True)        # loud on any mismatch (the reference hides it with strict=False)
            self.models.append(model.to(self.device).eval())
            self.num_folds += 1
        if self.num_folds == 0:
            print("[MIL] no folds found")
        else:
            init(self.device)
