"""``package_export``: contents.json + NNNN_shape.npy archive for the native runtime.

Format = what the reference produces for libZnicz (``Forward.package_export``
/root/reference/nn_units.py:152-161, spec visible in
/root/reference/libZnicz/tests/workflow_files/mnist.zip and
/root/reference/tests/functional/test_package_export.py:110-136)::

    contents.json = {"checksum": ..., "workflow": <class name>,
                     "units": [{"class": {"name", "uuid"},
                                "data": {attr: value | "@NNNN_shape"},
                                "links": [indices of the units fed by this one]}]}
    NNNN_<d0>x<d1>….npy  — one file per array, float16 (precision=16) or float32

Beyond the reference's three FC units, every forward unit with ``exports`` is packaged
(conv: kx, ky, n_kernels, padding, sliding; pooling: kx, ky, sliding; LRN: alpha, beta, k, n;
cutter: padding; dropout ratio), which is what ``veles.znicz_b200.native`` consumes.
"""
from __future__ import annotations

import hashlib
import io
import json
import tarfile
import time
import uuid
import zipfile

import numpy

_NS = uuid.UUID("5d6c8a52-8d55-4b0c-9a56-6d9f6e2b7a10")


def unit_uuid(unit):
    uid = getattr(type(unit), "__id__", None)
    if uid:
        return uid
    return str(uuid.uuid5(_NS, type(unit).__name__))


def _jsonable(v):
    if isinstance(v, (numpy.integer,)):
        return int(v)
    if isinstance(v, (numpy.floating,)):
        return float(v)
    if isinstance(v, (tuple, list)):
        return [_jsonable(x) for x in v]
    if isinstance(v, (bool, int, float, str)) or v is None:
        return v
    return str(v)


def collect(workflow, precision=32):
    """→ (contents dict, {file name: npy bytes})."""
    if precision not in (16, 32):
        raise ValueError("precision must be 16 or 32")
    dtype = numpy.float16 if precision == 16 else numpy.float32
    units = [u for u in getattr(workflow, "forwards", []) if hasattr(u, "package_export")]
    if not units:
        units = [u for u in workflow.units_in_dependency_order
                 if hasattr(u, "package_export")]
    index = {id(u): i for i, u in enumerate(units)}
    files = {}
    entries = []
    counter = 0
    for u in units:
        data = {}
        for name, value in sorted(u.package_export().items()):
            if isinstance(value, numpy.ndarray):
                arr = numpy.ascontiguousarray(value)
                if arr.dtype.kind == "f":
                    arr = arr.astype(dtype)
                fname = "%04d_%s" % (counter, "x".join(str(d) for d in arr.shape))
                counter += 1
                buf = io.BytesIO()
                numpy.save(buf, arr, allow_pickle=False)
                files[fname + ".npy"] = buf.getvalue()
                data[name] = "@" + fname
            else:
                data[name] = _jsonable(value)
        links = sorted(index[id(d)] for d in u.links_to if id(d) in index)
        entries.append({"class": {"name": type(u).__name__, "uuid": unit_uuid(u)},
                        "data": data, "links": links})
    h = hashlib.sha1()
    for e in entries:
        h.update(e["class"]["uuid"].encode())
    contents = {"checksum": "%s_%d" % (h.hexdigest(), len(entries)),
                "workflow": type(workflow).__name__, "units": entries}
    return contents, files


def package_export(workflow, file_name, archive_format="zip", precision=32):
    contents, files = collect(workflow, precision)
    blob = json.dumps(contents, indent=1, sort_keys=True).encode("utf-8")
    if archive_format == "zip":
        with zipfile.ZipFile(file_name, "w", zipfile.ZIP_DEFLATED) as z:
            z.writestr("contents.json", blob)
            for name, data in sorted(files.items()):
                z.writestr(name, data)
    elif archive_format == "tgz":
        with tarfile.open(file_name, "w:gz") as tar:
            def add(name, data):
                ti = tarfile.TarInfo(name)
                ti.size = len(data)
                ti.mtime = int(time.time())
                tar.addfile(ti, io.BytesIO(data))
            add("contents.json", blob)
            for name, data in sorted(files.items()):
                add(name, data)
    else:
        raise ValueError("archive_format must be 'zip' or 'tgz'")
    return file_name
