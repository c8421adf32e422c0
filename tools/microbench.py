"""Secondary BASELINE configs on one B200: MIL pooling (config 2), MIL train step, input stage (config 3), with the
reference's CPU path timed beside them.  Writes JSON lines (one per measurement) to stdout."""
import json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import teethrt
from teethrt import ops, preproc
from teethrt.modules import MILNet
from teethrt.train import MILTrainer
teethrt.init()
PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def gpu_time(fn, n=20, warm=3, l2flush=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        if l2flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return statistics.median(ts)


out = []
# ---- config 2: MIL gated-attention pooling alone
D, K, hid = 1280, 16, 128
Vw, Vb = torch.randn(hid, D, device="cuda") * D ** -0.5, torch.zeros(hid, device="cuda")
Uw, Ub = torch.randn(hid, D, device="cuda") * D ** -0.5, torch.zeros(hid, device="cuda")
ww, wb = torch.randn(hid, device="cuda") * hid ** -0.5, torch.zeros(1, device="cuda")
for B in (6, 64, 1024):
    H = torch.randn(B, K, D, device="cuda")
    t = gpu_time(lambda: ops.mil_attn_fwd(H, Vw, Vb, Uw, Ub, ww, wb), l2flush=False)
    alg = B * (K * D * 4 + D * 4 + K * 4) + 2 * hid * D * 4
    out.append({"what": "mil_attn_fwd", "B": B, "K": K, "D": D, "hid": hid, "us": t * 1e6, "bags_per_s": B / t,
                "algorithmic_GBps": alg / t / 1e9, "hbm_frac": alg / t / 1e9 / PK["hbm_gbs"], "gflops": B * 10.5e6 / t / 1e9})
if "--only-mil" in sys.argv:
    for o in out:
        print(json.dumps(o))
    sys.exit(0)
# ---- MIL train step (B=6 bags x 16 instances @224, B0 encoder)
torch.manual_seed(0)
m = MILNet().cuda()
tr = MILTrainer(m, lr=2e-4, t_max=1000, graph=True)
bags = torch.randn(6, 16, 3, 224, 224, device="cuda")
y = (torch.rand(6, device="cuda") < 0.6).float()
for _ in range(6):
    tr.step(bags, y)
torch.cuda.synchronize()
t = gpu_time(lambda: tr.step(bags, y), n=20, warm=2, l2flush=False)
out.append({"what": "mil_train_step", "bags": 6, "instances": 16, "img": 224, "ms": t * 1e3, "crops_per_s": 96 / t,
            "tflops": 96 * 3 * 0.769e9 / t / 1e12, "launches_per_step": tr.launches_per_step})
# ---- config 3: input stage on 1024^2 radiographs
import ref_preproc as P
import cv2
for n in (1, 64):
    imgs = torch.from_numpy(np.stack([P.image_set("radiograph", 1024, 1024, seed=i % 4) for i in range(n)])).cuda()
    dst = torch.empty_like(imgs)
    ws = torch.empty(teethrt.lib.trt_clahe_workspace_bytes(n), device="cuda", dtype=torch.uint8)
    tab = preproc.device_tables("cuda")
    from teethrt._lib import check, ptr, stream
    fn = lambda: check(teethrt.lib.trt_clahe_bgr_u8(ptr(imgs), ptr(dst), n, 1024, 1024, 3.0, ptr(tab), ptr(ws), ws.numel(), stream()))
    t = gpu_time(fn, l2flush=(n == 1))
    px = n * 1024 * 1024
    out.append({"what": "clahe_bgr_u8", "n": n, "us": t * 1e6, "Mpx_per_s": px / t / 1e6, "algorithmic_GBps": 6 * px / t / 1e9,
                "hbm_frac_6Bpx": 6 * px / t / 1e9 / PK["hbm_gbs"], "practical_GBps_9Bpx": 9 * px / t / 1e9})
    st = preproc.InputStage(n, 1024, 1024, size=224, dtype=torch.bfloat16)
    t = gpu_time(lambda: st(imgs, 0), l2flush=(n == 1))
    out.append({"what": "input_stage_clahe_resize_normalize", "n": n, "us": t * 1e6, "images_per_s": n / t,
                "algorithmic_GBps": n * 3.45e6 / t / 1e9})
img = P.image_set("radiograph", 1024, 1024)
for thr in (1, os.cpu_count()):
    cv2.setNumThreads(thr)
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); P.apply_clahe_cv2(img); ts.append(time.perf_counter() - t0)
    out.append({"what": "cpu_reference_apply_clahe", "threads": thr, "ms": statistics.median(ts) * 1e3,
                "Mpx_per_s": 1024 * 1024 / statistics.median(ts) / 1e6})
# ---- "next" rows of SURVEY 8 (f1-f4): device path (host buffers in, result on the host where the reference returns one)
# beside the reference's CPU path on this box's cores.  End-to-end wall clock, host->device copies included.
def wall(fn, n=10, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return statistics.median(ts)


cv2.setNumThreads(os.cpu_count())
tooth = P.tooth_image(1024, 1024, 3, 33.0)
tooth_dev = torch.from_numpy(tooth).cuda()
out.append({"what": "f1_deskew_1024", "device_resident_ms": wall(lambda: preproc.deskew(tooth_dev)) * 1e3,
            "host_in_host_out_ms": wall(lambda: preproc.deskew(tooth)) * 1e3, "cpu_reference_ms": wall(lambda: P.deskew_cv2(tooth)) * 1e3,
            "edges_us": gpu_time(lambda: preproc.canny(tooth_dev, want_edges=False), l2flush=False) * 1e6,
            "warp_us": gpu_time(lambda: preproc.warp_affine(tooth_dev, preproc.rotation_matrix_2d((512, 512), 33.0), (1024, 1024)), l2flush=False) * 1e6})
from PIL import Image
from torchvision import transforms
pil = Image.fromarray(tooth[..., ::-1].copy())
tfm = transforms.Compose([transforms.Resize(256, interpolation=transforms.InterpolationMode.BICUBIC), transforms.CenterCrop(224)])
rgb = np.asarray(pil)
out.append({"what": "f2_eval_resize_crop_1024_to_224", "host_in_device_out_ms": wall(lambda: preproc.resize_center_crop(rgb, 256, 224)) * 1e3,
            "kernels_us": gpu_time(lambda: preproc.resize_center_crop(tooth_dev, 256, 224), l2flush=False) * 1e6,
            "cpu_reference_pil_ms": wall(lambda: np.asarray(tfm(pil))) * 1e3})
import ref_calib as RC
import ref_stack as RS
from teethrt import calib, stack
z, yv, _ = RC.calib_cases()["large"]
zd, yd = torch.tensor(z).cuda(), torch.tensor(yv).cuda()
out.append({"what": "f3_calibrate_epoch_n10007", "device_ms": wall(lambda: calib.calibrate_epoch(zd, yd), n=5) * 1e3,
            "cpu_reference_ms": wall(lambda: RC.calibrate_epoch(z, yv), n=3, warm=1) * 1e3})
fr = RS.stream_frames(n=20000, n_test=4000)
oof = fr["tab_oof"].rename(columns={"prob": "a"}).merge(fr["mm_oof"].rename(columns={"prob": "b"}), on=["image_name", "y"])
X, yo = oof[["a", "b"]].values, oof["y"].values


def dev_stack():
    m = stack.LogisticMeta().fit(X, yo)
    p = m.predict_proba(X)[:, 1]
    return [stack.choose_threshold(yo, p, mode) for mode in RS.MODES]


def cpu_stack():
    m = RS.fit_meta(X, yo)
    p = m.predict_proba(X)[:, 1]
    return [RS.choose_threshold(yo, p, mode) for mode in RS.MODES]


out.append({"what": "f4_meta_fit_plus_5_threshold_modes_n20000", "device_ms": wall(dev_stack, n=5) * 1e3,
            "cpu_reference_ms": wall(cpu_stack, n=2, warm=1) * 1e3})
# f2 train transform: RandAugment operations on one 224x224 crop, device kernels beside Pillow (what timm calls) on the host
import ref_augment as RA
from teethrt import augment
crop = np.random.RandomState(0).randint(0, 256, (224, 224, 3), dtype=np.uint8)
crop_dev, crop_pil = torch.from_numpy(crop).cuda(), Image.fromarray(crop)
for name, args in (("Rotate", (27.0,)), ("ShearX", (0.27,)), ("SharpnessIncreasing", (1.81,)), ("ContrastIncreasing", (1.81,)), ("Equalize", ())):
    out.append({"what": "f2_randaugment_op_224", "op": name, "device_us": gpu_time(lambda: augment._OP_FN[name](crop_dev, *args), l2flush=False) * 1e6,
                "cpu_reference_pil_us": wall(lambda: RA.OPS[name](crop_pil, *args), n=20) * 1e3})
big = np.random.RandomState(1).randint(0, 256, (512, 512, 3), dtype=np.uint8)
tf = augment.TrainTransform(224, seed=0)
big_dev = torch.from_numpy(big).cuda()
out.append({"what": "f2_train_transform_512_to_224", "device_resident_ms": wall(lambda: tf(big_dev), n=30) * 1e3,
            "host_in_ms": wall(lambda: tf(big), n=30) * 1e3})
# f2, batched: one DataLoader batch (64 images, 512x512 uint8, the size src/preprocessing writes) per call, every stage one
# launch; beside (a) the per-image device path above looped over the batch and (b) the same transform on PIL images on the
# host (RandomResizedCrop bicubic + flip + the policy's Pillow ops + ToTensor/Normalize: what timm's transform does per image
# inside a DataLoader worker), one host thread
batch = np.random.RandomState(2).randint(0, 256, (64, 512, 512, 3), dtype=np.uint8)
batch_pin = torch.from_numpy(batch).pin_memory()
batch_dev = batch_pin.cuda()
btf = augment.BatchTrainTransform(224, seed=0)
plan = btf.sample(64, 512, 512)
t_sample = wall(lambda: btf.sample(64, 512, 512), n=20)
t_dev = wall(lambda: btf(batch_dev), n=20)
t_host = wall(lambda: btf(batch_pin), n=20)
t_loop = wall(lambda: [tf(batch_dev[i]) for i in range(64)], n=3, warm=1)
from torchvision.transforms import functional as TF


def pil_one(i):
    im = Image.fromarray(batch[i])
    top, left, ch, cw = (int(v) for v in plan.boxes[i])
    im = TF.resized_crop(im, top, left, ch, cw, [224, 224], interpolation=transforms.InterpolationMode.BICUBIC)
    if plan.flips[i]:
        im = TF.hflip(im)
    for layer in plan.layers:
        if layer[i] is not None:
            im = RA.OPS[layer[i][0]](im, *layer[i][1])
    return TF.normalize(TF.to_tensor(im), (0.485, 0.456, 0.406), (0.229, 0.224, 0.225))


t_pil = wall(lambda: [pil_one(i) for i in range(16)], n=3, warm=1) * 4
out.append({"what": "f2_train_transform_batch64_512_to_224", "device_resident_ms": t_dev * 1e3, "pinned_host_in_ms": t_host * 1e3,
            "host_sampling_ms": t_sample * 1e3, "images_per_s_device_resident": 64 / t_dev, "images_per_s_host_in": 64 / t_host,
            "per_image_device_loop_ms": t_loop * 1e3, "cpu_reference_pil_one_thread_ms": t_pil * 1e3,
            "h2d_bytes": int(batch.nbytes), "note": "sampling parity with timm unpinned; pixels Pillow-exact"})
for o in out:
    print(json.dumps(o))