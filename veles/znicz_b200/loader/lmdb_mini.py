"""Minimal LMDB environment reader / bulk writer (pure Python).

Caffe datasets are LMDB environments (/root/reference/loader/loader_lmdb.py:147-169);
the ``lmdb`` binding is not part of the target image, so the on-disk B+tree is read
directly. Layout (LMDB 0.9 ``mdb.c``, 64-bit little endian):

  page header (16 B): pgno u64 | pad u16 | flags u16 | lower u16, upper u16  (or
                      overflow page count u32 for P_OVERFLOW)
  flags: P_BRANCH 0x01, P_LEAF 0x02, P_OVERFLOW 0x04, P_META 0x08
  meta pages 0 and 1 (the one with the larger txnid is current), after the header:
      magic u32 0xBEEFC0DE | version u32 | address u64 | mapsize u64 |
      2 x MDB_db {pad u32 (page size in FREE_DBI) | flags u16 | depth u16 | branch u64 |
                  leaf u64 | overflow u64 | entries u64 | root u64} | last_pg u64 | txnid u64
  node: lo u16 | hi u16 | flags u16 | ksize u16 | key | data
      leaf:   data size = lo | hi << 16; F_BIGDATA (0x01): data = u64 overflow pgno
      branch: child pgno = lo | hi << 16 | flags << 32

``Environment`` / ``Cursor`` mimic the subset of the ``lmdb`` API the loader needs
(``first``, ``next``, ``key``, ``value``, ``get``, ``stat``). ``write_environment`` bulk
loads sorted (key, value) pairs into a fresh environment (tests, dataset preparation).
"""
from __future__ import annotations

import builtins
import mmap
import os
import struct

_open_file = builtins.open

P_BRANCH, P_LEAF, P_OVERFLOW, P_META = 0x01, 0x02, 0x04, 0x08
F_BIGDATA = 0x01
MAGIC, VERSION = 0xBEEFC0DE, 1
PAGEHDRSZ, NODESIZE = 16, 8
P_INVALID = 0xFFFFFFFFFFFFFFFF
_META = struct.Struct("<IIQQ")
_DB = struct.Struct("<IHHQQQQQ")


class Error(Exception):
    pass


class Environment(object):
    def __init__(self, path, readonly=True, **_ignored):
        self.path = os.path.join(path, "data.mdb") if os.path.isdir(path) else path
        self._file = _open_file(self.path, "rb")
        size = os.fstat(self._file.fileno()).st_size
        if size < 2 * 512:
            raise Error("%s is too small to be an LMDB environment" % self.path)
        self._map = mmap.mmap(self._file.fileno(), 0, access=mmap.ACCESS_READ)
        self.buf = memoryview(self._map)
        best = None
        # the page size is recorded in meta 0; meta 1 sits one page later
        m0 = self._read_meta(0)
        self.psize = m0["psize"]
        for off in (0, self.psize):
            m = self._read_meta(off)
            if best is None or m["txnid"] > best["txnid"]:
                best = m
        self.meta = best

    def _read_meta(self, off):
        flags = struct.unpack_from("<H", self.buf, off + 10)[0]
        if not flags & P_META:
            raise Error("page at %d is not a meta page" % off)
        magic, version, _addr, mapsize = _META.unpack_from(self.buf, off + PAGEHDRSZ)
        if magic != MAGIC:
            raise Error("bad LMDB magic %#x" % magic)
        if version != VERSION:
            raise Error("unsupported LMDB data version %d" % version)
        p = off + PAGEHDRSZ + _META.size
        free = _DB.unpack_from(self.buf, p)
        main = _DB.unpack_from(self.buf, p + _DB.size)
        last_pg, txnid = struct.unpack_from("<QQ", self.buf, p + 2 * _DB.size)
        return {"psize": free[0], "mapsize": mapsize, "depth": main[2], "branch": main[3],
                "leaf": main[4], "overflow": main[5], "entries": main[6], "root": main[7],
                "last_pg": last_pg, "txnid": txnid}

    def stat(self):
        m = self.meta
        return {"psize": self.psize, "depth": m["depth"], "branch_pages": m["branch"],
                "leaf_pages": m["leaf"], "overflow_pages": m["overflow"],
                "entries": m["entries"]}

    # -- pages / nodes ---------------------------------------------------------------------
    def page(self, pgno):
        off = pgno * self.psize
        flags, lower, upper = struct.unpack_from("<HHH", self.buf, off + 10)
        return off, flags, (lower - PAGEHDRSZ) >> 1

    def node(self, page_off, i):
        ptr = struct.unpack_from("<H", self.buf, page_off + PAGEHDRSZ + 2 * i)[0]
        lo, hi, flags, ksize = struct.unpack_from("<HHHH", self.buf, page_off + ptr)
        return page_off + ptr + NODESIZE, lo, hi, flags, ksize

    def leaf_value(self, data_off, lo, hi, flags, ksize):
        size = lo | (hi << 16)
        p = data_off + ksize
        if flags & F_BIGDATA:
            pgno = struct.unpack_from("<Q", self.buf, p)[0]
            p = pgno * self.psize + PAGEHDRSZ
        return self.buf[p:p + size]

    # -- lmdb-like API -------------------------------------------------------------------------
    def begin(self, **_ignored):
        return self

    def cursor(self):
        return Cursor(self)

    def get(self, key, default=None):
        c = Cursor(self)
        return c.value() if c.set_key(key) else default

    def close(self):
        self.buf.release()
        self._map.close()
        self._file.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def open(path, **kwargs):          # noqa: A001  (mirrors ``lmdb.open``)
    return Environment(path, **kwargs)


class Cursor(object):
    def __init__(self, env):
        self.env = env
        self.stack = []            # [(page_off, index, n_keys)] root → leaf
        self._valid = False

    def _descend(self, pgno, leftmost=True, key=None):
        env = self.env
        while True:
            off, flags, n = env.page(pgno)
            if flags & P_LEAF:
                return off, n
            if not flags & P_BRANCH:
                raise Error("unexpected page flags %#x" % flags)
            idx = 0
            if key is not None:           # last child whose separator key <= key
                lo_, hi_ = 1, n - 1
                while lo_ <= hi_:
                    mid = (lo_ + hi_) >> 1
                    doff, _l, _h, _f, ks = env.node(off, mid)
                    if bytes(env.buf[doff:doff + ks]) <= key:
                        idx = mid
                        lo_ = mid + 1
                    else:
                        hi_ = mid - 1
            elif not leftmost:
                idx = n - 1
            self.stack.append((off, idx, n))
            _d, lo, hi, fl, _k = env.node(off, idx)
            pgno = lo | (hi << 16) | (fl << 32)

    def first(self):
        self.stack = []
        root = self.env.meta["root"]
        if root == P_INVALID or not self.env.meta["entries"]:
            self._valid = False
            return False
        off, n = self._descend(root)
        self.stack.append((off, 0, n))
        self._valid = n > 0
        return self._valid

    def next(self):
        if not self._valid:
            return False
        off, i, n = self.stack[-1]
        if i + 1 < n:
            self.stack[-1] = (off, i + 1, n)
            return True
        # climb until a branch has a right sibling
        self.stack.pop()
        while self.stack:
            boff, bi, bn = self.stack.pop()
            if bi + 1 < bn:
                self.stack.append((boff, bi + 1, bn))
                _d, lo, hi, fl, _k = self.env.node(boff, bi + 1)
                loff, ln = self._descend(lo | (hi << 16) | (fl << 32))
                self.stack.append((loff, 0, ln))
                return True
        self._valid = False
        return False

    def set_key(self, key):
        self.stack = []
        root = self.env.meta["root"]
        if root == P_INVALID:
            self._valid = False
            return False
        off, n = self._descend(root, key=bytes(key))
        env = self.env
        lo_, hi_ = 0, n - 1
        while lo_ <= hi_:
            mid = (lo_ + hi_) >> 1
            doff, _l, _h, _f, ks = env.node(off, mid)
            k = bytes(env.buf[doff:doff + ks])
            if k == key:
                self.stack.append((off, mid, n))
                self._valid = True
                return True
            if k < key:
                lo_ = mid + 1
            else:
                hi_ = mid - 1
        self._valid = False
        return False

    def get(self, key, default=None):
        return self.value() if self.set_key(key) else default

    def key(self):
        off, i, _n = self.stack[-1]
        doff, _lo, _hi, _fl, ks = self.env.node(off, i)
        return bytes(self.env.buf[doff:doff + ks])

    def value(self):
        off, i, _n = self.stack[-1]
        return bytes(self.env.leaf_value(*self.env.node(off, i)))

    def item(self):
        return self.key(), self.value()

    def __iter__(self):
        ok = self.first()
        while ok:
            yield self.item()
            ok = self.next()


# -- bulk writer ------------------------------------------------------------------------------
def _even(n):
    return (n + 1) & ~1


def write_environment(path, items, psize=4096, mapsize=1 << 30):
    """Create ``path/data.mdb`` (+ empty lock file) holding ``items`` ((key, value) bytes
    pairs; sorted here). Values that do not fit a node go to overflow pages."""
    items = sorted((bytes(k), bytes(v)) for k, v in items)
    os.makedirs(path, exist_ok=True)
    nodemax = (((psize - PAGEHDRSZ) // 2) & ~1) - 2
    pages = [None, None]                   # page images; metas patched at the end
    n_over = n_leaf = n_branch = 0

    def new_page(flags):
        pages.append(None)
        return len(pages) - 1

    def build_page(pgno, flags, nodes):
        """nodes: list of bytes (already NODESIZE header + key + data)."""
        img = bytearray(psize)
        upper = psize
        ptrs = []
        for nd in nodes:
            upper -= _even(len(nd))
            img[upper:upper + len(nd)] = nd
            ptrs.append(upper)
        lower = PAGEHDRSZ + 2 * len(nodes)
        assert lower <= upper
        struct.pack_into("<QHHHH", img, 0, pgno, 0, flags, lower, upper)
        struct.pack_into("<%dH" % len(ptrs), img, PAGEHDRSZ, *ptrs)
        pages[pgno] = bytes(img)

    def leaf_node(key, value):
        nonlocal n_over
        if NODESIZE + len(key) + len(value) > nodemax:
            npg = (PAGEHDRSZ + len(value) + psize - 1) // psize
            first = len(pages)
            blob = bytearray(npg * psize)
            struct.pack_into("<QHHI", blob, 0, first, 0, P_OVERFLOW, npg)
            blob[PAGEHDRSZ:PAGEHDRSZ + len(value)] = value
            for i in range(npg):
                pages.append(bytes(blob[i * psize:(i + 1) * psize]))
            n_over += npg
            return struct.pack("<HHHH", len(value) & 0xFFFF, len(value) >> 16, F_BIGDATA,
                               len(key)) + key + struct.pack("<Q", first)
        return struct.pack("<HHHH", len(value) & 0xFFFF, len(value) >> 16, 0,
                           len(key)) + key + value

    def pack_level(entries, flags):
        """entries: [(first_key, node_bytes)] → [(first_key, pgno)] of the pages created."""
        out, cur, cur_key, used = [], [], None, PAGEHDRSZ
        for key, nd in entries:
            need = _even(len(nd)) + 2
            if cur and used + need > psize:
                pg = new_page(flags)
                build_page(pg, flags, cur)
                out.append((cur_key, pg))
                cur, used = [], PAGEHDRSZ
            if not cur:
                cur_key = key
                if flags == P_BRANCH:       # the first key of a branch page is implicit
                    nd = nd[:6] + struct.pack("<H", 0) + b""
                    need = _even(len(nd)) + 2
            cur.append(nd)
            used += need
        if cur:
            pg = new_page(flags)
            build_page(pg, flags, cur)
            out.append((cur_key, pg))
        return out

    root, depth = P_INVALID, 0
    if items:
        level = pack_level([(k, leaf_node(k, v)) for k, v in items], P_LEAF)
        n_leaf = len(level)
        depth = 1
        while len(level) > 1:
            entries = [(k, struct.pack("<HHHH", pg & 0xFFFF, (pg >> 16) & 0xFFFF,
                                       (pg >> 32) & 0xFFFF, len(k)) + k)
                       for k, pg in level]
            level = pack_level(entries, P_BRANCH)
            n_branch += len(level)
            depth += 1
        root = level[0][1]
    last_pg = len(pages) - 1
    for i, txnid in ((0, 0), (1, 1)):
        img = bytearray(psize)
        struct.pack_into("<QHHHH", img, 0, i, 0, P_META, 0, 0)
        p = PAGEHDRSZ
        _META.pack_into(img, p, MAGIC, VERSION, 0, mapsize)
        p += _META.size
        _DB.pack_into(img, p, psize, 0, 0, 0, 0, 0, 0, P_INVALID)
        p += _DB.size
        if txnid:
            _DB.pack_into(img, p, 0, 0, depth, n_branch, n_leaf, n_over, len(items), root)
        else:
            _DB.pack_into(img, p, 0, 0, 0, 0, 0, 0, 0, P_INVALID)
        p += _DB.size
        struct.pack_into("<QQ", img, p, last_pg if txnid else 1, txnid)
        pages[i] = bytes(img)
    with _open_file(os.path.join(path, "data.mdb"), "wb") as f:
        for pg in pages:
            f.write(pg)
    _open_file(os.path.join(path, "lock.mdb"), "wb").close()
    return len(items)
