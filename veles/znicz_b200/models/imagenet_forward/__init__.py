"""ImagenetAE localisation pipeline: sliding bounding boxes -> classifier -> merged detections.

Capability parity with /root/reference/tests/research/ImagenetAE/imagenet_forward/ (2.1 k LoC):

* ``bbox``       box geometry, non-maximum suppression and the probability-weighted box merging
                 (forward_bbox.py:43-613) - vectorised over all boxes instead of python pair loops;
* ``loader``     ``ForwardLoaderBbox``: every candidate box of every image, rotated by a set of
                 angles and mirrored, cut out, scaled to the network aperture and alpha-blended
                 over the training mean (forward_loader.py:59-477) - own affine sampler, no OpenCV;
* ``merge``      ``MergeBboxes``: per-image accumulation of class probabilities per box and the two
                 decision modes "merge" and "final" (imagenet_forward.py:62-235);
* ``writer``     ``ResultWriter`` (forward_json.py:46-119) and the DET / CLS-LOC text converters,
                 json merging and raw box extraction (json2txt.py, merge_json.py,
                 raw_bbox_extract.py);
* ``workflow``   ``ImagenetForward``: loader -> mean/disp normaliser -> trained forward units ->
                 merge -> writer, run twice (merge stage, then final stage on its own output;
                 imagenet_forward.py:237-388), and ``shard_range`` - the image range of one
                 rank (distribute_forward.py:40-85 used ssh fan-out; here: torchrun ranks).
"""
from .bbox import (BBox, overlap_area, overlap_ratio, has_inclusion, is_small, nms_detections,  # noqa: F401
                   merge_to_one, merge_by_probs, merge_by_dict, remove_inner,
                   postprocess_same_label)
