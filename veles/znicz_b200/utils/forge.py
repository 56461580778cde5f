"""Model packaging (``manifest.json``): the reference ships every sample as
``<name>.py`` + ``<name>_config.py`` + ``manifest.json`` for its VelesForge hub
(/root/reference/samples/Wine/manifest.json). Here the manifests live in
``models/manifests/*.json``; ``pack()`` writes the same kind of self-describing archive
(manifest + workflow module), ``unpack()`` reads it back, ``list_models()`` enumerates."""
from __future__ import annotations

import io
import json
import os
import tarfile

_HERE = os.path.dirname(os.path.abspath(__file__))
MODELS_DIR = os.path.join(os.path.dirname(_HERE), "models")
MANIFESTS_DIR = os.path.join(MODELS_DIR, "manifests")
REQUIRED = ("name", "workflow", "configuration", "short_description", "long_description",
            "requires", "author")


def list_models():
    return sorted(f[:-5] for f in os.listdir(MANIFESTS_DIR) if f.endswith(".json"))


def load_manifest(name):
    with open(os.path.join(MANIFESTS_DIR, name + ".json")) as f:
        m = json.load(f)
    missing = [k for k in REQUIRED if k not in m]
    if missing:
        raise ValueError("manifest %s lacks %s" % (name, missing))
    if not os.path.isfile(os.path.join(MODELS_DIR, m["workflow"])):
        raise ValueError("manifest %s: workflow file %s not found" % (name, m["workflow"]))
    return m


def pack(name, path):
    """``<path>`` (tar.gz): manifest.json + the files the manifest names."""
    m = load_manifest(name)
    with tarfile.open(path, "w:gz") as tar:
        blob = json.dumps(m, indent=4, sort_keys=True).encode()
        info = tarfile.TarInfo("manifest.json")
        info.size = len(blob)
        tar.addfile(info, io.BytesIO(blob))
        for fn in sorted(set([m["workflow"]] + list(m.get("files", [])))):
            full = os.path.join(MODELS_DIR, fn)
            if os.path.isfile(full):
                tar.add(full, arcname=fn)
    return path


def unpack(path, directory):
    with tarfile.open(path, "r:gz") as tar:
        names = tar.getnames()
        if "manifest.json" not in names:
            raise ValueError("not a model package: no manifest.json")
        for member in tar.getmembers():       # flat archives only
            if os.path.basename(member.name) != member.name:
                raise ValueError("unexpected path in package: %s" % member.name)
        tar.extractall(directory)
    with open(os.path.join(directory, "manifest.json")) as f:
        return json.load(f)
