"""One-time conversion: the reference's DBoW2 vocabulary voc/point_voc_L4.bin (a Boost binary archive of
TemplatedVocabulary<FSuperpoint>: include/bow/database.h:35-55) -> weights/point_voc_L4.afw (AIRFEW01 container, see tools/make_weights.py).

Run in the authoring container only (needs /root/reference/voc/point_voc_L4.bin):
    python -m tools.make_voc

Archive layout (boost::archive::binary_oarchive, library version 17, little endian), as written by the serialize() overloads:
  header | class header (5 B) | m_k i32 | m_L i32 | m_weighting i32 | m_scoring i32 | vector<Node> class header (5 B) | count u64 | item version u32 |
  per node: [first node: class header 5 B] object id u32 | id u32 | weight f64 | children: count u64 + u32[count] | parent u32 | descriptor f32[256] | word_id u32
  (m_words, a vector of pointers into m_nodes, follows and is not needed: every leaf carries its word id).
The same parser exists in C++ (airslam_b200/csrc/bow.cu: load_boost_vocabulary) for drop-in loading of the original file."""
import os
import struct

import numpy as np

from .make_weights import OUT, write_container

SRC = "/root/reference/voc/point_voc_L4.bin"


def parse_boost_vocabulary(path):
    d = open(path, "rb").read()
    n, = struct.unpack_from("<Q", d, 0)
    assert d[8:8 + n] == b"serialization::archive"
    off = 8 + n + 2 + 4 + 4 + 5
    k, L, weighting, scoring = struct.unpack_from("<4i", d, off)
    off += 16 + 5
    cnt, = struct.unpack_from("<Q", d, off)
    off += 12
    children = np.full((cnt, k), -1, np.int32)
    desc = np.zeros((cnt, 256), np.float32)
    word_id = np.zeros(cnt, np.int32)
    weight = np.zeros(cnt, np.float64)
    parent = np.zeros(cnt, np.int32)
    for i in range(cnt):
        if i == 0:
            off += 5
        _, nid = struct.unpack_from("<II", d, off)
        off += 8
        assert nid == i
        weight[i], = struct.unpack_from("<d", d, off)
        off += 8
        nc, = struct.unpack_from("<Q", d, off)
        off += 8
        assert nc <= k
        children[i, :nc] = struct.unpack_from("<%dI" % nc, d, off)
        off += 4 * nc
        parent[i], = struct.unpack_from("<I", d, off)
        off += 4
        desc[i] = np.frombuffer(d, np.float32, 256, off)
        off += 1024
        word_id[i], = struct.unpack_from("<I", d, off)
        off += 4
    return dict(k=k, L=L, weighting=weighting, scoring=scoring, children=children, desc=desc, word_id=word_id, weight=weight, parent=parent)


def main():
    v = parse_boost_vocabulary(SRC)
    meta = np.array([v["k"], v["L"], v["weighting"], v["scoring"]], np.int32)
    write_container(os.path.join(OUT, "point_voc_L4.afw"),
                    [("voc.meta", meta), ("voc.children", v["children"]), ("voc.desc", v["desc"].reshape(-1)), ("voc.word_id", v["word_id"]),
                     ("voc.weight_f64_bits", v["weight"].view(np.int32).reshape(-1, 2))])


if __name__ == "__main__":
    main()
