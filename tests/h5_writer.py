"""Test infrastructure: a minimal pure-Python HDF5 *writer* producing files shaped like the ones CLAM's
``save_hdf5`` writes through h5py (utils/file_utils.py:16-35 of the reference): superblock version 0, a
symbol-table root group (v1 group B-tree + local heap + one SNOD), version-1 object headers, and per dataset
either chunked storage indexed by a (possibly multi-level) v1 chunk B-tree -- chunk shape ``(1, ...)``,
resizable along axis 0, chunks of the datasets interleaved in the file the way batch-wise appends leave them --
or contiguous / compact storage.  Written from the HDF5 File Format Specification (version 1.1 structures);
neither h5py nor libhdf5 exists in this image, so this writer is the only producer of chunked fixtures and the
reader's chunked path is pinned by it, while the contiguous path is pinned by a file libhdf5 itself wrote
(scipy's MATLAB-7.3 test file, see tests/test_h5.py).
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


class _File:
    def __init__(self, user_block: int = 0, superblock_bytes: int = 96):
        self.buf = bytearray(user_block + superblock_bytes + (-superblock_bytes % 8))
        self.base = user_block

    def alloc(self, data: bytes, align: int = 8) -> int:
        """Appends data; returns its address relative to the base address."""
        while (len(self.buf) - self.base) % align:
            self.buf.append(0)
        addr = len(self.buf) - self.base
        self.buf += data
        return addr

    def patch(self, addr: int, data: bytes) -> None:
        self.buf[self.base + addr:self.base + addr + len(data)] = data


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages: Sequence[bytes], split_at: Optional[int] = None, f: Optional[_File] = None) -> bytes:
    """Version-1 object header.  With split_at, messages[split_at:] go to a continuation block (needs f)."""
    if split_at is None:
        body = b"".join(messages)
        return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body
    tail = b"".join(messages[split_at:])
    cont_addr = f.alloc(tail)
    head = list(messages[:split_at]) + [_message(0x0010, struct.pack("<QQ", cont_addr, len(tail)))]
    body = b"".join(head)
    return struct.pack("<BxHII4x", 1, len(messages) + 1, 1, len(body)) + body


def _datatype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f":
        exp_bits, man_bits = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[dt.itemsize]
        bias = (1 << (exp_bits - 1)) - 1
        head = struct.pack("<BBBBI", 0x11, 0x20, dt.itemsize * 8 - 1, 0, dt.itemsize)
        return head + struct.pack("<HHBBBBI", 0, dt.itemsize * 8, man_bits, exp_bits, 0, man_bits, bias)
    if dt.kind in "iu":
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0, 0, 0, dt.itemsize)
        return head + struct.pack("<HH", 0, dt.itemsize * 8)
    raise TypeError(dt)


def _dataspace(shape: Tuple[int, ...], resizable: bool) -> bytes:
    body = struct.pack("<BBB5x", 1, len(shape), 1 if resizable else 0)
    body += b"".join(struct.pack("<Q", int(d)) for d in shape)
    if resizable:
        body += struct.pack("<Q", UNDEF) + b"".join(struct.pack("<Q", int(d)) for d in shape[1:])
    return body


def _chunk_btree(f: _File, entries: List[Tuple[Tuple[int, ...], int]], chunk_bytes: int, rank: int, fanout: int,
                 end_key: Tuple[int, ...]) -> int:
    """entries: (chunk offset tuple, address), ascending.  Returns the root node's address."""
    def key(off: Tuple[int, ...], size: int) -> bytes:
        return struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", int(o)) for o in off) + struct.pack("<Q", 0)

    level = 0
    nodes = [(off, addr) for off, addr in entries]
    while True:
        parents = []
        groups = [nodes[i:i + fanout] for i in range(0, len(nodes), fanout)] or [[]]
        addrs = []
        for g in groups:
            body = b""
            for off, addr in g:
                body += key(off, chunk_bytes) + struct.pack("<Q", addr)
            body += key(end_key, 0)
            node = b"TREE" + struct.pack("<BBH", 1, level, len(g)) + struct.pack("<QQ", UNDEF, UNDEF) + body
            full = 24 + (2 * fanout) * (8 + 8 * (rank + 1) + 8) + 8 + 8 * (rank + 1)   # libhdf5 allocates full nodes
            node += b"\0" * max(0, full - len(node))
            a = f.alloc(node)
            addrs.append(a)
            parents.append((g[0][0] if g else end_key, a))
        for i, a in enumerate(addrs):   # sibling pointers
            left = addrs[i - 1] if i > 0 else UNDEF
            right = addrs[i + 1] if i + 1 < len(addrs) else UNDEF
            f.patch(a + 8, struct.pack("<QQ", left, right))
        if len(parents) == 1:
            return parents[0][1]
        nodes = parents
        level += 1


def write_h5(path: str, datasets: Dict[str, np.ndarray], layout: str = "chunked", batch: int = 512,
             chunk_fanout: int = 64, user_block: int = 0, superblock_version: int = 0,
             split_headers: bool = False, missing_chunks: Optional[Dict[str, Sequence[int]]] = None,
             chunk_rows: int = 1) -> None:
    """Writes ``datasets`` (name -> C-contiguous ndarray, rank >= 1) into the root group of a new file.

    layout: "chunked" (chunk shape (chunk_rows, *shape[1:]), resizable axis 0, appended ``batch`` rows at a time in
    round-robin over the datasets, like CLAM's save_hdf5 loop), "contiguous" or "compact".
    missing_chunks: per dataset, chunk indices left unallocated (they read back as the fill value 0).
    """
    f = _File(user_block, 100 if superblock_version == 1 else 96)
    names = sorted(datasets)          # SNOD entries must be sorted by name
    arrays = {n: np.ascontiguousarray(datasets[n]) for n in names}
    missing = {n: set(missing_chunks.get(n, ())) if missing_chunks else set() for n in names}

    # local heap: empty string at offset 0, then the names
    heap_data = bytearray(8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap_data)
        heap_data += _pad8(n.encode() + b"\0")
    heap_size = max(len(heap_data) + 16, 88)
    free_off = len(heap_data)
    heap_data += struct.pack("<QQ", 1, heap_size - free_off)      # free block: next (1 = none), size
    heap_data += b"\0" * (heap_size - len(heap_data))
    heap_addr = f.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, heap_size, free_off, 0))
    heap_data_addr = f.alloc(bytes(heap_data))
    f.patch(heap_addr + 24, struct.pack("<Q", heap_data_addr))

    # raw data
    data_addr: Dict[str, int] = {}
    chunk_lists: Dict[str, List[Tuple[Tuple[int, ...], int]]] = {n: [] for n in names}
    if layout == "chunked":
        longest = max(a.shape[0] for a in arrays.values()) if names else 0
        for lo in range(0, longest, batch):
            for n in names:
                a = arrays[n]
                for r in range(lo, min(lo + batch, a.shape[0]), chunk_rows):
                    if r // chunk_rows in missing[n]:
                        continue
                    piece = a[r:r + chunk_rows]
                    if piece.shape[0] < chunk_rows:    # edge chunk is stored full-size
                        pad = np.zeros((chunk_rows - piece.shape[0],) + a.shape[1:], dtype=a.dtype)
                        piece = np.concatenate([piece, pad])
                    addr = f.alloc(piece.tobytes(), align=1)
                    chunk_lists[n].append(((r,) + (0,) * (a.ndim - 1), addr))
    elif layout == "contiguous":
        for n in names:
            data_addr[n] = f.alloc(arrays[n].tobytes()) if arrays[n].size else UNDEF

    # dataset object headers
    ohdr_addr = {}
    for n in names:
        a = arrays[n]
        msgs = [_message(0x0001, _dataspace(a.shape, layout == "chunked")), _message(0x0003, _datatype(a.dtype), 1),
                _message(0x0005, struct.pack("<BBBB", 2, 3 if layout == "chunked" else 2, 2, 0))]
        if layout == "chunked":
            cdims = (chunk_rows,) + a.shape[1:]
            cbytes = int(np.prod(cdims)) * a.dtype.itemsize
            end_key = (((a.shape[0] + chunk_rows - 1) // chunk_rows) * chunk_rows,) + (0,) * (a.ndim - 1)
            root = _chunk_btree(f, chunk_lists[n], cbytes, a.ndim, chunk_fanout, end_key) if chunk_lists[n] else UNDEF
            body = struct.pack("<BBB", 3, 2, a.ndim + 1) + struct.pack("<Q", root)
            body += b"".join(struct.pack("<I", int(d)) for d in cdims) + struct.pack("<I", a.dtype.itemsize)
        elif layout == "contiguous":
            body = struct.pack("<BB", 3, 1) + struct.pack("<QQ", data_addr[n], a.nbytes)
        elif layout == "compact":
            assert a.nbytes < 60000, "compact datasets live inside the object header"
            body = struct.pack("<BBH", 3, 0, a.nbytes) + a.tobytes()
        else:
            raise ValueError(layout)
        msgs.append(_message(0x0008, body))
        msgs.append(_message(0x0012, struct.pack("<B3xI", 1, 1700000000)))
        if split_headers:
            msgs.insert(2, _message(0x0000, b"\0" * 24))     # a NIL message, as libhdf5 leaves after edits
            hdr = _object_header(msgs, split_at=2, f=f)
        else:
            hdr = _object_header(msgs)
        ohdr_addr[n] = f.alloc(hdr)

    # symbol table node + group B-tree + root object header
    snod = b"SNOD" + struct.pack("<BxH", 1, len(names))
    for n in names:
        snod += struct.pack("<QQII16x", name_off[n], ohdr_addr[n], 0, 0)
    snod += b"\0" * (8 + 8 * 40 - len(snod))           # room for 2K = 8 entries
    assert len(names) <= 8
    snod_addr = f.alloc(snod)
    last = name_off[names[-1]] if names else 0
    tree = b"TREE" + struct.pack("<BBH", 0, 0, 1) + struct.pack("<QQ", UNDEF, UNDEF)
    tree += struct.pack("<QQQ", 0, snod_addr, last)
    tree += b"\0" * (24 + 32 * 16 + 8 - len(tree))     # internal K = 16
    btree_addr = f.alloc(tree)
    root_ohdr = f.alloc(_object_header([_message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))]))

    eof = len(f.buf) - f.base
    sb = SIG + struct.pack("<BBBBBBBB", superblock_version, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0)
    if superblock_version == 1:
        sb += struct.pack("<HH", 32, 0)
    sb += struct.pack("<QQQQ", f.base if user_block else 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_ohdr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
    assert len(sb) == (100 if superblock_version == 1 else 96)
    f.buf[f.base:f.base + len(sb)] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(f.buf))
