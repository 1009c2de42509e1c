"""TEST INFRASTRUCTURE: a minimal HDF5 *writer* producing the file format the HDF5 library's default (earliest) settings
give -- superblock version 0, old-style groups (symbol-table message, v1 B-tree with 8-entry symbol-table nodes, local
heap, links sorted by name), contiguous little-endian float32 datasets -- so that tests/ can hand
text_to_speech_b200.h5lite / convert.from_keras_weights_h5 a `.weights.h5` laid out the way Keras 3 lays out the
reference's WaveGlow (keras is not installable here). Written from the HDF5 file-format specification; the READER is
additionally checked against a file produced by the real HDF5 library (tests/test_convert.py)."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)           # superblock v0 with 8-byte offsets, filled in at the end

    def alloc(self, data):
        off = len(self.buf)
        self.buf += _pad8(bytes(data))
        return off

    def dataset(self, arr):
        arr = np.ascontiguousarray(arr, dtype="<f4")
        data_addr = self.alloc(arr.tobytes()) if arr.size else UNDEF
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        dtype = struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        layout = struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes)
        msgs = b""
        for mtype, body in ((1, space), (3, dtype), (8, layout)):
            body = _pad8(body)
            msgs += struct.pack("<HHB3x", mtype, len(body), 0) + body
        return self.alloc(struct.pack("<BBHII4x", 1, 0, 3, 1, len(msgs)) + msgs)

    def group(self, children):
        """children: {name: ('group', header, btree, heap) | ('dataset', header)} -> (header, btree, heap)"""
        names = sorted(children, key=lambda s: s.encode())          # the library orders links by strcmp
        heap = bytearray(8)                                          # offset 0: the empty string
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode() + b"\0")
        heap_data = self.alloc(heap)
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), UNDEF, heap_data))
        snods, keys = [], [0]
        for s in range(0, len(names), 2 * LEAF_K):
            part = names[s:s + 2 * LEAF_K]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for n in part:
                c = children[n]
                if c[0] == "group":
                    body += struct.pack("<QQII", offs[n], c[1], 1, 0) + struct.pack("<QQ", c[2], c[3])
                else:
                    body += struct.pack("<QQII", offs[n], c[1], 0, 0) + b"\0" * 16
            body += b"\0" * (40 * (2 * LEAF_K - len(part)))
            snods.append(self.alloc(body))
            keys.append(offs[part[-1]])
        if len(snods) > 2 * INTERNAL_K:
            raise ValueError("too many links for a single-level B-tree")
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
        for i, a in enumerate(snods):
            node += struct.pack("<QQ", keys[i], a)
        node += struct.pack("<Q", keys[len(snods)])
        node += b"\0" * (24 + (2 * INTERNAL_K) * 16 + 8 - len(node))
        btree = self.alloc(node)
        msg = struct.pack("<HHB3x", 0x11, 16, 0) + struct.pack("<QQ", btree, heap_addr)
        header = self.alloc(struct.pack("<BBHII4x", 1, 0, 1, 1, len(msg)) + msg)
        return header, btree, heap_addr

    def finish(self, root):
        header, btree, heap = root
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, header, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_h5(path, datasets):
    """datasets: {'a/b/c': ndarray} -> an HDF5 file with one group per path component."""
    tree = {}
    for k, v in datasets.items():
        node = tree
        parts = k.strip("/").split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = np.asarray(v)
    w = _Writer()

    def build(node):
        children = {}
        for name, child in node.items():
            if isinstance(child, dict):
                children[name] = ("group",) + build(child)
            else:
                children[name] = ("dataset", w.dataset(child))
        return w.group(children)

    with open(path, "wb") as f:
        f.write(w.finish(build(tree)))


def keras3_waveglow_layout(hp, weights, fused=False):
    """{h5 dataset path: array} for the reference's architectures.WaveGlow as Keras 3's saving_lib lays it out
    (attribute names, list elements named by class in snake case with a running suffix, variables as vars/<i>)."""
    sfx = lambda base, i: base if i == 0 else f"{base}_{i}"
    out = {"upsample/vars/0": weights["upsample/kernel"], "upsample/vars/1": weights["upsample/bias"]}
    for k in range(hp.n_flows):
        b, src = f"blocks/{sfx('waveglow_block', k)}", f"block-{k}"
        out[f"convinv/{sfx('invertible1x1_conv', k)}/conv/vars/0"] = weights[f"invertible_conv-{k}/conv/kernel"]
        for attr, name in (("start", "start_conv"), ("end", "end_conv")):
            out[f"{b}/{attr}/vars/0"], out[f"{b}/{attr}/vars/1"] = weights[f"{src}/{name}/kernel"], weights[f"{src}/{name}/bias"]
        for attr, name in (("in_layers", "in_conv"), ("res_skip_layers", "res_skip_conv")) + ((() if fused else (("cond_layers", "cond_layer"),))):
            for i in range(hp.n_layers):
                out[f"{b}/{attr}/{sfx('conv1d', i)}/vars/0"] = weights[f"{src}/{name}-{i}/kernel"]
                out[f"{b}/{attr}/{sfx('conv1d', i)}/vars/1"] = weights[f"{src}/{name}-{i}/bias"]
        if fused:
            out[f"{b}/cond_layer/vars/0"] = np.concatenate([weights[f"{src}/cond_layer-{i}/kernel"] for i in range(hp.n_layers)], axis=2)
            out[f"{b}/cond_layer/vars/1"] = np.concatenate([weights[f"{src}/cond_layer-{i}/bias"] for i in range(hp.n_layers)])
    return out
