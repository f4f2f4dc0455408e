"""Fixtures of the partitioner's host logic, produced by the COMPILED REFERENCE (oracle/_ref/libifref_cpu.so:
InferenceEngine::LoadDeviceGroups on a real .ini read by the reference's ConfigData, NetworkBuilder::SplitGpuLayers).
Run in the authoring container: python tests/golden/make_golden_partition.py -> tests/golden/partition.json"""
import ctypes as C
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import orc  # noqa: E402

DEVICE_STRINGS = [
    "0", "1", "0;1", "0,1,2,3", "0&1", "0|1|2|3", "0&1&2&3&4&5&6&7", "0&1;2&3", "0&1&2&3;4&5&6&7", "0|1;2|3;4|5;6|7",
    " 0 & 1 ;\t2 | 3 ; ", "0;;1", "0;1;", ";0;1", "3&1;0&2", "gpu1&2", "0&1;2", "0;1&2", "0&&1", "&", "0&1;&", "",
    "7", "0 1", "0;1;2;3;4;5;6;7", "0&1,2&3", "0|1,2&3;4|5", "-1", "1.5&2", "0x1&2",
]
SPLITS = [(0, 32, 1), (0, 32, 2), (0, 32, 3), (0, 32, 4), (0, 32, 8), (0, 80, 8), (0, 80, 3), (0, 12, 5), (0, 5, 4), (0, 2, 4),
          (0, 1, 1), (0, 1, 3), (0, 7, 2), (0, 0, 2), (0, 126, 8), (0, 40, 6)]


def ref_lib():
    lib = C.CDLL(orc.ref_cpu_path())
    lib.ifref_load_device_groups.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.ifref_split_gpu_layers.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    return lib


def ref_groups(lib, text, tmpdir):
    """the string goes through the reference's own .ini reader, as in a deployment"""
    path = os.path.join(tmpdir, "g.ini")
    with open(path, "w") as f:
        f.write("[transformer_engine]\ndevices = %s\nmodels = x\n" % text)
    ids, sizes = (C.c_int32 * 256)(), (C.c_int32 * 64)()
    n = lib.ifref_load_device_groups(path.encode(), b"transformer_engine", ids, 256, sizes, 64)
    if n < 0:
        return {"error": n}
    out, k = [], 0
    for g in range(n):
        out.append([ids[k + i] for i in range(sizes[g])])
        k += sizes[g]
    return {"groups": out}


def ref_split(lib, a, b, n):
    r = (C.c_int32 * (3 * 64))()
    k = lib.ifref_split_gpu_layers(a, b, n, r, 64)
    return [[r[3 * i], r[3 * i + 1], r[3 * i + 2]] for i in range(k)]


def main():
    lib = ref_lib()
    with tempfile.TemporaryDirectory() as d:
        grammar = [{"devices": s, **ref_groups(lib, s, d)} for s in DEVICE_STRINGS]
    splits = [{"start": a, "end": b, "groups": n, "ranges": ref_split(lib, a, b, n)} for a, b, n in SPLITS]
    with open(os.path.join(HERE, "partition.json"), "w") as f:
        json.dump({"made_by": "tests/golden/make_golden_partition.py from oracle/_ref/libifref_cpu.so (the compiled reference)",
                   "grammar": grammar, "splits": splits}, f, indent=1)
    print(json.dumps(grammar)[:1500])
    print(json.dumps(splits)[:800])


if __name__ == "__main__":
    main()
