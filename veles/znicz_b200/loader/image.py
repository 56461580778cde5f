"""Image-file loaders (``veles.loader.image`` / ``file_image`` / ``fullbatch_image``
equivalents used by the reference samples: Kanji
/root/reference/samples/Kanji/kanji_config.py:55-70, Lines
/root/reference/samples/Lines/lines_config.py:65-70, YaleFaces
/root/reference/samples/YaleFaces/yale_faces_config.py:63-74).

Registered names:
  ``full_batch_file_image``                 labels from ``label_regexp`` on the file name
  ``full_batch_auto_label_file_image``      label = name of the containing directory
  ``full_batch_auto_label_file_image_mse``  + per-label target images (``target_paths``)
  ``auto_label_file_image``                 streaming variant: decodes per minibatch

Decoding uses OpenCV when present (PIL otherwise). Options: ``color_space`` (RGB, GRAY,
HSV, YCR_CB, ...), ``scale`` (float | (w, h)) with ``scale_maintain_aspect_ratio`` and
``background_color`` padding, ``crop`` ((w, h) centre crop), ``add_sobel`` (extra gradient
magnitude channel), ``mirror`` (True = add mirrored copies of TRAIN samples),
``rotations`` (list of angles in radians, extra rotated TRAIN copies), ``file_subtypes``
(MIME sub-types), ``ignored_files`` / ``included_files`` (regexps).
"""
from __future__ import annotations

import mimetypes
import os
import re

import numpy

from ..core.memory import Array
from .base import Loader, LoaderError, TEST, VALID, TRAIN
from .fullbatch import FullBatchLoader, FullBatchLoaderMSE

_EXTRA_TYPES = {".pgm": "x-portable-graymap", ".ppm": "x-portable-pixmap",
                ".pbm": "x-portable-bitmap", ".jpg": "jpeg", ".jpe": "jpeg",
                ".jpeg": "jpeg", ".png": "png", ".bmp": "bmp", ".tif": "tiff",
                ".tiff": "tiff", ".gif": "gif", ".webp": "webp"}


def image_subtype(path):
    ext = os.path.splitext(path)[1].lower()
    if ext in _EXTRA_TYPES:
        return _EXTRA_TYPES[ext]
    mt = mimetypes.guess_type(path)[0]
    if mt and mt.startswith("image/"):
        return mt.split("/", 1)[1]
    return None


def _cv2():
    try:
        import cv2
        return cv2
    except ImportError:      # pragma: no cover
        return None


def read_image(path, color_space="RGB"):
    """→ uint8 HWC array (HW1 for GRAY) in ``color_space``."""
    cv2 = _cv2()
    if cv2 is not None:
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img is None:
            raise LoaderError("cannot decode image %s" % path)
        if img.dtype != numpy.uint8:
            img = (img.astype(numpy.float32) * (255.0 / max(float(img.max()), 1.0))) \
                .astype(numpy.uint8)
        if img.ndim == 2:
            src = "GRAY"
        elif img.shape[2] == 4:
            img, src = cv2.cvtColor(img, cv2.COLOR_BGRA2BGR), "BGR"
        else:
            src = "BGR"
        if src != color_space:
            if src == "BGR" and color_space not in ("RGB", "GRAY"):
                img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
                src = "RGB"
            if src == "GRAY" and color_space not in ("RGB", "BGR"):
                img = cv2.cvtColor(img, cv2.COLOR_GRAY2RGB)
                src = "RGB"
            if src != color_space:
                img = cv2.cvtColor(img, getattr(cv2, "COLOR_%s2%s" % (src, color_space)))
    else:                    # pragma: no cover
        from PIL import Image
        im = Image.open(path)
        im = im.convert("L" if color_space == "GRAY" else "RGB")
        img = numpy.asarray(im)
    if img.ndim == 2:
        img = img[:, :, None]
    return numpy.ascontiguousarray(img)


def fit_image(img, scale=1.0, maintain_aspect=False, background_color=None, crop=None):
    """scale → (pad to the exact target with ``background_color``) → centre crop."""
    cv2 = _cv2()
    h, w = img.shape[:2]
    if isinstance(scale, (tuple, list)):
        tw, th = int(scale[0]), int(scale[1])
    else:
        tw, th = int(round(w * scale)), int(round(h * scale))
    if (tw, th) != (w, h):
        if maintain_aspect:
            k = min(tw / w, th / h)
            nw, nh = max(1, int(round(w * k))), max(1, int(round(h * k)))
        else:
            nw, nh = tw, th
        c = img.shape[2]
        if cv2 is not None:
            res = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_AREA
                             if nw * nh < w * h else cv2.INTER_CUBIC)
        else:                # pragma: no cover
            from PIL import Image
            res = numpy.asarray(Image.fromarray(img.squeeze()).resize((nw, nh)))
        res = res.reshape(nh, nw, c)
        if (nw, nh) != (tw, th):
            bg = numpy.zeros(c, numpy.uint8)
            if background_color is not None:
                bg[:] = numpy.asarray(background_color, numpy.uint8)[:c]
            canvas = numpy.empty((th, tw, c), numpy.uint8)
            canvas[:] = bg
            y0, x0 = (th - nh) // 2, (tw - nw) // 2
            canvas[y0:y0 + nh, x0:x0 + nw] = res
            res = canvas
        img = res
    if crop is not None:
        cw, ch = int(crop[0]), int(crop[1])
        h, w = img.shape[:2]
        y0, x0 = max(0, (h - ch) // 2), max(0, (w - cw) // 2)
        img = img[y0:y0 + ch, x0:x0 + cw]
    return numpy.ascontiguousarray(img)


def sobel_channel(img):
    cv2 = _cv2()
    gray = img.mean(axis=2).astype(numpy.float32)
    if cv2 is not None:
        gx = cv2.Sobel(gray, cv2.CV_32F, 1, 0, ksize=3)
        gy = cv2.Sobel(gray, cv2.CV_32F, 0, 1, ksize=3)
    else:                    # pragma: no cover
        gy, gx = numpy.gradient(gray)
    mag = numpy.hypot(gx, gy)
    mx = mag.max()
    if mx:
        mag *= 255.0 / mx
    return mag.astype(numpy.uint8)[:, :, None]


def rotate_image(img, angle, background_color=None):
    cv2 = _cv2()
    if not angle:
        return img
    h, w = img.shape[:2]
    m = cv2.getRotationMatrix2D((w / 2.0, h / 2.0), numpy.degrees(angle), 1.0)
    bg = tuple(int(v) for v in (background_color or (0,) * img.shape[2]))
    out = cv2.warpAffine(img, m, (w, h), borderValue=bg)
    return out.reshape(h, w, img.shape[2])


class ImageOptionsMixin(object):
    """Parsing of the shared image kwargs + file discovery."""

    def _init_image_options(self, kwargs):
        self.color_space = kwargs.get("color_space", "RGB")
        self.scale = kwargs.get("scale", 1.0)
        self.scale_maintain_aspect_ratio = kwargs.get("scale_maintain_aspect_ratio", False)
        self.crop = kwargs.get("crop")
        self.background_color = kwargs.get("background_color")
        self.add_sobel = kwargs.get("add_sobel", False)
        self.mirror = kwargs.get("mirror", False)
        self.rotations = tuple(kwargs.get("rotations", (0.0,)))
        self.file_subtypes = [s.lower() for s in kwargs.get("file_subtypes", ())]
        self.ignored_files = [re.compile(p) for p in kwargs.get("ignored_files", ())]
        self.included_files = [re.compile(p) for p in kwargs.get("included_files", (".*",))]
        self.test_paths = list(kwargs.get("test_paths", ()))
        self.validation_paths = list(kwargs.get("validation_paths", ()))
        self.train_paths = list(kwargs.get("train_paths", ()))
        self.label_regexp = re.compile(kwargs["label_regexp"]) \
            if kwargs.get("label_regexp") else None

    def is_valid_filename(self, path):
        name = os.path.basename(path)
        if any(p.match(name) for p in self.ignored_files):
            return False
        if not any(p.match(name) for p in self.included_files):
            return False
        st = image_subtype(path)
        if st is None:
            return False
        return not self.file_subtypes or st in self.file_subtypes

    def scan_files(self, paths):
        keys = []
        for base in paths:
            if os.path.isfile(base):
                if self.is_valid_filename(base):
                    keys.append(base)
                continue
            if not os.path.isdir(base):
                raise LoaderError("image path %s does not exist" % base)
            for d, _dirs, files in sorted(os.walk(base, followlinks=True)):
                keys.extend(os.path.join(d, f) for f in sorted(files)
                            if self.is_valid_filename(os.path.join(d, f)))
        return keys

    def get_image_label(self, key):
        if self.label_regexp is not None:
            m = self.label_regexp.search(os.path.basename(key))
            if m is None:
                raise LoaderError("label_regexp does not match %s" % key)
            return m.group(1) if m.groups() else m.group(0)
        return os.path.basename(os.path.dirname(key))

    def decode(self, key):
        img = read_image(key, self.color_space)
        img = fit_image(img, self.scale, self.scale_maintain_aspect_ratio,
                        self.background_color, self.crop)
        if self.add_sobel:
            img = numpy.concatenate([img, sobel_channel(img)], axis=2)
        return img

    def variants(self, img, train):
        """Augmented copies made at load time (TRAIN only)."""
        out = [img]
        if not train:
            return out
        for a in self.rotations:
            if a:
                out.append(rotate_image(img, a, self.background_color))
        if self.mirror is True:
            out.extend([v[:, ::-1].copy() for v in list(out)])
        return out


class FullBatchFileImageLoader(FullBatchLoader, ImageOptionsMixin):
    """All images decoded once into ``original_data`` (uint8 → normalised float)."""
    MAPPING = "full_batch_file_image"

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self._init_image_options(kwargs)
        self.class_keys = [[], [], []]

    def _labelled(self):
        return True

    def load_data(self):
        chunks, labels = [], []
        shape = None
        for cls, paths in ((TEST, self.test_paths), (VALID, self.validation_paths),
                           (TRAIN, self.train_paths)):
            keys = self.scan_files(paths)
            self.class_keys[cls] = keys
            n = 0
            for k in keys:
                img = self.decode(k)
                if shape is None:
                    shape = img.shape
                elif img.shape != shape:
                    raise LoaderError(
                        "%s has shape %s, expected %s (set scale/crop)" % (k, img.shape, shape))
                vs = self.variants(img, cls == TRAIN)
                chunks.extend(vs)
                if self._labelled():
                    labels.extend([self.get_image_label(k)] * len(vs))
                n += len(vs)
            self.class_lengths[cls] = n
        if not chunks:
            raise LoaderError("no images found in %s" % (
                self.test_paths + self.validation_paths + self.train_paths))
        self.original_data.reset(numpy.stack(chunks))
        self.original_labels = labels
        self.info("Loaded %d images of shape %s (%d labels)", len(chunks), shape,
                  len(set(labels)))


class FullBatchAutoLabelFileImageLoader(FullBatchFileImageLoader):
    MAPPING = "full_batch_auto_label_file_image"


class FullBatchAutoLabelFileImageLoaderMSE(FullBatchLoaderMSE, ImageOptionsMixin):
    """Targets are images: ``target_paths`` holds one image per label (file stem or its
    directory = label); every sample's target is its label's image (``class_targets``)."""
    MAPPING = "full_batch_auto_label_file_image_mse"

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self._init_image_options(kwargs)
        self.target_paths = list(kwargs.get("target_paths", ()))
        self.targets_shape_hint = kwargs.get("targets_shape")
        self.class_keys = [[], [], []]

    load_images = FullBatchFileImageLoader.load_data
    _labelled = FullBatchFileImageLoader._labelled

    def load_data(self):
        self.load_images()
        tkeys = self.scan_files(self.target_paths)
        if not tkeys:
            raise LoaderError("no target images in %s" % self.target_paths)
        by_label = {}
        for k in tkeys:
            stem = os.path.splitext(os.path.basename(k))[0]
            img = read_image(k, self.color_space)
            if self.targets_shape_hint is not None:
                th, tw = self.targets_shape_hint[:2]
                img = fit_image(img, (tw, th))
            by_label[stem] = img
            by_label.setdefault(os.path.basename(os.path.dirname(k)), img)
        uniq = sorted(set(self.original_labels))
        missing = [l for l in uniq if l not in by_label]
        if missing:
            raise LoaderError("no target image for labels %s" % missing[:5])
        ct = numpy.stack([by_label[l] for l in uniq]).astype(numpy.float32)
        if ct.shape[-1] == 1:
            ct = ct[..., 0]
        self.class_targets.reset(ct)
        lm = {l: i for i, l in enumerate(uniq)}
        self.labels_mapping = lm
        self.reversed_labels_mapping = uniq
        idx = numpy.fromiter((lm[l] for l in self.original_labels), dtype=numpy.int64,
                             count=len(self.original_labels))
        self.original_targets.reset(ct[idx])


class ImageLoaderBase(Loader):
    """Streaming image loader in terms of three callbacks (the reference's ``IImageLoader``:
    /root/reference/loader/loader_lmdb.py:76-104, loader_stl.py:96-116):

      ``get_keys(class_index)``   → list of opaque keys of that class
      ``get_image_data(key)``     → HWC array
      ``get_image_label(key)``    → raw label

    Only keys are kept in memory; each minibatch is decoded on demand. Normalisation
    statistics come from a bounded random subset of TRAIN (``analysis_samples``)."""
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.analysis_samples = kwargs.get("analysis_samples", 1024)
        self.mirror = kwargs.get("mirror", False)
        self.keys = []
        self.key_labels = []
        self.sample_shape = None

    def get_keys(self, index):
        raise NotImplementedError

    def get_image_data(self, key):
        raise NotImplementedError

    def get_image_label(self, key):
        raise NotImplementedError

    def _data_loaded(self):
        return bool(self.keys)

    def load_data(self):
        self.keys, raw = [], []
        for cls in (TEST, VALID, TRAIN):
            ks = list(self.get_keys(cls))
            self.class_lengths[cls] = len(ks)
            self.keys.extend(ks)
            raw.extend(self.get_image_label(k) for k in ks)
        if not self.keys:
            raise LoaderError("no images found")
        uniq = sorted(set(raw))
        self.labels_mapping = {l: i for i, l in enumerate(uniq)}
        self.reversed_labels_mapping = uniq
        self.key_labels = numpy.array([self.labels_mapping[l] for l in raw], numpy.int32)
        self.sample_shape = self.get_image_data(self.keys[-1]).shape

    @property
    def has_labels(self):
        return True

    def create_minibatch_data(self):
        from ..core.accelerated_units import host_dtype
        shape = (self.max_minibatch_size,) + tuple(self.sample_shape)
        self.minibatch_data.reset(numpy.zeros(shape, dtype=host_dtype()))
        self.minibatch_labels.reset(numpy.zeros(self.max_minibatch_size, numpy.int32))
        if self.on_cuda:
            from ..ops.nn_units import torch_act_dtype
            self.minibatch_data.dev_dtype = torch_act_dtype()

    def analyze_dataset(self):
        norm = self.normalizer
        if self.normalization_type == "none" or norm.is_initialized:
            return
        start = self.class_end_offsets[VALID]
        n = self.class_lengths[TRAIN] or self.total_samples
        if not self.class_lengths[TRAIN]:
            start = 0
        pick = numpy.random.RandomState(1).permutation(n)[:self.analysis_samples] + start
        sample = numpy.stack([self.get_image_data(self.keys[i]) for i in pick])
        norm.analyze(sample.astype(numpy.float32))

    def fill_minibatch(self):
        n = self.minibatch_size
        idx = self.minibatch_indices.mem[:n]
        self.minibatch_data.map_invalidate()
        self.minibatch_labels.map_invalidate()
        md, ml = self.minibatch_data.mem, self.minibatch_labels.mem
        raw = numpy.stack([self.get_image_data(self.keys[i]) for i in idx]).astype(md.dtype)
        if self.mirror == "random" and self.minibatch_class == TRAIN:
            flip = self.prng.randint(0, 2, n).astype(bool) if hasattr(self.prng, "randint") \
                else numpy.random.randint(0, 2, n).astype(bool)
            raw[flip] = raw[flip][:, :, ::-1]
        if self.normalization_type != "none":
            raw = self.normalizer.normalize(raw)
        md[:n] = raw
        md[n:] = 0
        ml[:n] = self.key_labels[idx]
        ml[n:] = -1


class FullBatchImageLoaderBase(FullBatchLoader):
    """Full-batch counterpart of ``ImageLoaderBase`` (same three callbacks)."""
    hide_from_registry = True

    get_keys = ImageLoaderBase.get_keys
    get_image_data = ImageLoaderBase.get_image_data
    get_image_label = ImageLoaderBase.get_image_label

    def load_data(self):
        chunks, labels = [], []
        for cls in (TEST, VALID, TRAIN):
            ks = list(self.get_keys(cls))
            self.class_lengths[cls] = len(ks)
            chunks.extend(self.get_image_data(k) for k in ks)
            labels.extend(self.get_image_label(k) for k in ks)
        if not chunks:
            raise LoaderError("no images found")
        self.original_data.reset(numpy.stack(chunks))
        self.original_labels = labels


class AutoLabelFileImageLoader(ImageOptionsMixin, ImageLoaderBase):
    """Streaming file loader: label = directory name (or ``label_regexp``)."""
    MAPPING = "auto_label_file_image"

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self._init_image_options(kwargs)

    def get_keys(self, index):
        return self.scan_files((self.test_paths, self.validation_paths,
                                self.train_paths)[index])

    def get_image_data(self, key):
        return self.decode(key)
