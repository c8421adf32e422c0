"""ORACLE: `timm.data.create_transform` as the reference calls it
(experiments/multimodal_v1/train_mm_joint_dualtask.py:72-93, ui/gradio_app/infer_mm.py:12-17).

Eval: Resize(floor(S/0.875), bicubic) -> CenterCrop(S) -> ToTensor -> Normalize.  The training
variant (RandAugment etc., SURVEY.md §8 row f2 = "next") is approximated with torchvision parts;
it is outside the measured hot path (BASELINE uses synthetic tensors).
"""
import math

from torchvision import transforms as T

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def create_transform(input_size, is_training=False, interpolation="bicubic", mean=IMAGENET_DEFAULT_MEAN,
                     std=IMAGENET_DEFAULT_STD, auto_augment=None, re_prob=0.0, re_mode="const", re_count=1,
                     crop_pct=0.875, **kwargs):
    size = input_size if isinstance(input_size, int) else input_size[-1]
    interp = {"bicubic": T.InterpolationMode.BICUBIC, "bilinear": T.InterpolationMode.BILINEAR}[interpolation]
    if not is_training:
        return T.Compose([
            T.Resize(int(math.floor(size / crop_pct)), interpolation=interp),
            T.CenterCrop(size),
            T.ToTensor(),
            T.Normalize(mean, std),
        ])
    tf = [T.RandomResizedCrop(size, scale=(0.08, 1.0), ratio=(3 / 4, 4 / 3), interpolation=interp),
          T.RandomHorizontalFlip(0.5)]
    if auto_augment:
        tf.append(T.RandAugment(num_ops=2, magnitude=9))
    tf += [T.ToTensor(), T.Normalize(mean, std)]
    if re_prob > 0:
        tf.append(T.RandomErasing(p=re_prob, value="random" if re_mode == "pixel" else 0))
    return T.Compose(tf)
