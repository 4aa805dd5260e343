"""``data_dict.npz`` of the multi-view reconstruction data (DSS/utils/dataset.py:30-211, written by
scripts/create_mvr_data_from_mesh.py:170-255): the cloud in object coordinates and one 4x4 world-to-view matrix per
image, plus the ground-truth image and mask files next to it (dataset.py:36-101: every file of ``<data_dir>/image`` /
``<data_dir>/mask`` with the configured extension, sorted; rgb -> (3,H,W) float32 in [0,1], mask -> (1,H,W) float 0/1),
read with PIL (the reference uses imageio, which this image lacks).  ``pinned_batch`` stages a batch of views in pinned
host memory for the step's host->device copy.  Dense depth maps (OpenEXR) are not read.  Host-side and cold."""
import os

import numpy as np
import torch

from ..core.camera import FoVPerspectiveCameras
from ..core.cloud import PointClouds3D

__all__ = ["MVRData", "decompose_to_R_and_t"]


def decompose_to_R_and_t(transform_mat, row_major=True):
    """4x4 transform -> R (..,3,3), t (..,3) in the row-vector convention X_view = X_world R + t
    (DSS/utils/mathHelper.py:163-174)."""
    if transform_mat.shape[-2:] != (4, 4):
        raise ValueError("Expecting batches of 4x4 matrices")
    if not row_major:
        transform_mat = transform_mat.transpose(-2, -1)
    return transform_mat[..., :3, :3], transform_mat[..., -1, :3]


class MVRData:
    """points / normals / colors and per-image ``camera_mat`` of a ``data_dict.npz``.

    ``data[i]`` -> ``{"camera_mat": (4,4) f32 [, "img.rgb": (3,H,W), "img.mask": (1,H,W)]}`` like
    ``MVRDataset.__getitem__`` (dataset.py:171-211); images are only present when given to the constructor."""

    def __init__(self, data_dir_or_file, data_dict="data_dict.npz", images=None, masks=None, n_imgs=None,
                 img_folder="image", mask_folder="mask", img_extension="png", mask_extension="png", load_images=True):
        path = data_dir_or_file
        data_dir = None
        if os.path.isdir(path):
            data_dir = path
            path = os.path.join(path, data_dict)
        if load_images and data_dir is not None and images is None and masks is None:
            # dataset.py:41-57: the files of the two folders with the configured extension, sorted by name
            def listing(folder, ext):
                d = os.path.join(data_dir, folder)
                if not os.path.isdir(d):
                    return None
                return [os.path.join(d, f) for f in sorted(os.listdir(d)) if os.path.splitext(f)[1].lower()[1:] == ext]
            self.image_files, self.mask_files = listing(img_folder, img_extension), listing(mask_folder, mask_extension)
            if self.image_files:
                images = self.load_all_images()
            if self.mask_files:
                masks = self.load_all_masks()
        self.data_dict = np.load(path, allow_pickle=True)
        if "camera_mat" not in self.data_dict:
            raise ValueError("data_dict must contain camera_mat!")
        self.camera_mat = np.asarray(self.data_dict["camera_mat"], dtype=np.float32)
        if self.camera_mat.ndim != 3 or self.camera_mat.shape[1:] != (4, 4):
            raise ValueError("camera_mat must be (n_images, 4, 4), got %s" % (self.camera_mat.shape,))
        for name, arr in (("images", images), ("masks", masks)):
            if arr is not None and len(arr) != len(self.camera_mat):
                raise ValueError("Found unequal number of %s and camera matrices! (%d, %d)"
                                 % (name, len(arr), len(self.camera_mat)))
        self.images, self.masks = images, masks
        self.n_imgs = len(self.camera_mat) if n_imgs is None else min(int(n_imgs), len(self.camera_mat))

    def __len__(self):
        return self.n_imgs

    def load_all_images(self):
        """dataset.py:88-93: (3,H,W) float32 tensors in [0,1], alpha dropped."""
        from PIL import Image
        out = []
        for path in self.image_files:
            rgb = np.asarray(Image.open(path).convert("RGB"), dtype=np.float32) / 255.0
            out.append(torch.from_numpy(np.ascontiguousarray(rgb.transpose(2, 0, 1))))
        return out

    def load_all_masks(self):
        """dataset.py:95-101: 8-bit grey -> bool -> (1,H,W) float32 0/1."""
        from PIL import Image
        out = []
        for path in self.mask_files:
            m = np.asarray(Image.open(path).convert("L")).astype(bool)[None]
            out.append(torch.from_numpy(np.ascontiguousarray(m)).float())
        return out

    @property
    def resolution(self):
        return None if self.images is None else tuple(self.images[0].shape[1:])

    def pinned_batch(self, indices):
        """Views `indices` stacked in PINNED host memory, ready for one non-blocking copy per tensor:
        {"camera_mat": (B,4,4) [, "img.rgb": (B,3,H,W), "img.mask": (B,1,H,W)]} (what a training step uploads)."""
        idx = [int(i) % len(self) for i in indices]
        pin = lambda t: t.pin_memory() if torch.cuda.is_available() else t
        out = {"camera_mat": pin(torch.from_numpy(np.stack([self.camera_mat[i] for i in idx])))}
        if self.images is not None:
            out["img.rgb"] = pin(torch.stack([torch.as_tensor(self.images[i]) for i in idx]))
        if self.masks is not None:
            out["img.mask"] = pin(torch.stack([torch.as_tensor(self.masks[i]) for i in idx]))
        return out

    def __getitem__(self, idx):
        idx = idx % len(self)
        out = {"camera_mat": self.camera_mat[idx].copy()}
        if self.images is not None:
            out["img.rgb"] = self.images[idx]
        if self.masks is not None:
            out["img.mask"] = self.masks[idx]
        lights = self.data_dict["lights_%d" % idx] if ("lights_%d" % idx) in self.data_dict else None
        if lights is not None:
            props = lights.item() if hasattr(lights, "item") and lights.dtype == object else None
            if isinstance(props, dict):
                out["lights"] = {k: np.array(v, dtype=np.float32)[0] for k, v in props.items()
                                 if isinstance(v, (list, np.ndarray))}
        return out

    def get_pointclouds(self) -> PointClouds3D:
        """points, normals and colours in object coordinates (dataset.py:107-135, the stored-cloud branch)."""
        t = lambda k: torch.tensor(np.asarray(self.data_dict[k]), dtype=torch.float32)
        colors = t("colors") if "colors" in self.data_dict else torch.ones_like(t("points"))
        return PointClouds3D([t("points")], [t("normals")], [colors])

    def get_cameras(self, camera_mat=None, **camera_params):
        """``FoVPerspectiveCameras`` with R, T taken from ``camera_mat`` ((n,4,4) or (4,4); default: all images)
        (dataset.py:155-165; the stored ``cameras_params`` are used when present)."""
        params = dict(camera_params)
        if not params and "cameras_params" in self.data_dict:
            stored = self.data_dict["cameras_params"]
            stored = stored.item() if hasattr(stored, "item") else stored
            if isinstance(stored, dict):
                params = {k: v for k, v in stored.items() if k in ("znear", "zfar", "aspect_ratio", "fov", "degrees")}
        m = torch.as_tensor(self.camera_mat[:len(self)] if camera_mat is None else np.asarray(camera_mat), dtype=torch.float32)
        if m.dim() == 2:
            m = m[None]
        R, T = decompose_to_R_and_t(m)
        return FoVPerspectiveCameras(R=R.contiguous(), T=T.contiguous(), **params)
