"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol that
include/sa_engine.h declares, the helper entry points follow the reference's rules, and the
product refuses to run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from spectral_analyzer_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sa_engine.h")).read()
    return sorted(set(re.findall(r"SA_API\s+[\w\s\*]+?\b(sa_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _capi.lib()
    names = declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(L, n), "libsa_engine.so does not export " + n


def test_bytes_per_iq_matches_global_java():
    L = _capi.lib()      # S/sigmf/Global.java:67-79
    assert [L.sa_bytes_per_iq(i) for i in range(6)] == [8, 4, 2, 2, 16, 0]


@pytest.mark.parametrize("s,exp", [("cf32_le", (0, 0)), ("cf32_be", (0, 1)), ("ci16_le", (1, 0)), ("ci16_be", (1, 1)),
                                   ("cu8", (2, 1)), ("ci8", (3, 1)), ("cf64_le", (4, 0)), ("ci16", (1, 1))])
def test_parse_datatype(s, exp):
    assert _capi.parse_datatype(s) == exp


def test_parse_datatype_unknown_is_error():
    with pytest.raises(_capi.EngineError) as ei:
        _capi.parse_datatype("ri16_le")          # mono WAV: no decode branch in the reference
    assert ei.value.code == 2


def test_params_defaults_are_the_reference_mode():
    p = _capi.default_params()
    assert p.struct_size == C.sizeof(_capi.SpectrogramParams)
    assert (p.window, p.nfft, p.hop, p.db_mode, p.eof_fill_db) == (0, 1024, 1024, 0, -150.0)
    assert (p.min_db, p.max_db) == (-160.0, -30.0)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = _capi.lib().sa_engine_create(0, C.byref(h))
    assert rc == 5 and not h.value
    assert b"no CPU fallback" in _capi.lib().sa_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "spectral_analyzer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), f + " mentions the oracle"


def test_only_tests_smoke_and_the_bench_baseline_use_the_oracle():
    """Outside tests/ and oracle/ itself, only bench.py (cpu_baseline / --impl reference legs) and
    __graft_entry__.py (smoke) may import it: tools, the other bench scripts, headers and profiles helpers must not."""
    allowed = {"bench.py", "__graft_entry__.py"}
    offenders = []
    for dirpath, dirs, files in os.walk(ROOT):
        rel = os.path.relpath(dirpath, ROOT)
        top = rel.split(os.sep)[0]
        if top in ("tests", "oracle", ".git", "gpurun_out", "baseline", ".pytest_cache", ".hypothesis") or "__pycache__" in rel:
            dirs[:] = []
            continue
        for f in files:
            if not f.endswith((".py", ".hpp", ".h", ".java", ".cu", ".cuh")) or (rel == "." and f in allowed):
                continue
            txt = open(os.path.join(dirpath, f), errors="ignore").read()
            if "import oracle" in txt or "from oracle" in txt or "c_oracle" in txt or "np_oracle" in txt or "sa_oracle" in txt:
                offenders.append(os.path.join(rel, f))
    assert not offenders, offenders


def test_python_binding_covers_the_header():
    """_capi.py is the 1:1 mirror of the Panama binding: every function the header declares has its argument
    types spelled out there (no implicit int marshalling of 64-bit sizes or pointers)."""
    L = _capi.lib()
    no_args = {"sa_last_error", "sa_version"}
    for n in declared_symbols():
        if n in no_args:
            continue
        assert getattr(L, n).argtypes is not None, n + " has no argtypes in _capi.py"


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Argument errors of the canvas / packer / series calls are reported before any device work."""
    L = _capi.lib()
    assert L.sa_iq_pack(None, None, None, 0, 0, None) == 1            # engine is NULL
    assert b"engine is NULL" in L.sa_last_error()
    p = _capi.default_params()
    assert L.sa_render_canvas(None, None, 0, C.byref(p), 4, 4, 1, 0, None) == 1


def test_java_binding_source_matches_the_header():
    """The Panama binding cannot be compiled here (no JDK), so its structure is checked as text: every symbol it looks
    up is declared in include/sa_engine.h, and its StructLayout of sa_spectrogram_params lists the header's fields in
    the header's order with matching widths (the same order the ctypes mirror uses)."""
    java = open(os.path.join(ROOT, "java", "net", "kcundercover", "spectral_analyzer", "services",
                             "NativeSpectralEngine.java")).read()
    header = open(os.path.join(ROOT, "include", "sa_engine.h")).read()
    looked_up = set(re.findall(r'fn\("(sa_\w+)"', java))
    assert len(looked_up) >= 12 and looked_up <= set(declared_symbols()), looked_up - set(declared_symbols())
    body = re.search(r"typedef struct sa_spectrogram_params \{(.*?)\} sa_spectrogram_params;", header, re.S).group(1)
    c_fields = re.findall(r"^\s*(uint32_t|int32_t|uint64_t|double)\s+(\w+);", body, re.M)
    layout = re.search(r"PARAMS = MemoryLayout\.structLayout\((.*?)\);", java, re.S).group(1)
    j_fields = re.findall(r'(JAVA_INT|JAVA_LONG|JAVA_DOUBLE)\.withName\("(\w+)"\)', layout)
    width = {"uint32_t": "JAVA_INT", "int32_t": "JAVA_INT", "uint64_t": "JAVA_LONG", "double": "JAVA_DOUBLE"}
    assert [(width[t], n) for t, n in c_fields] == j_fields
    assert [n for _, n in c_fields] == [n for n, _ in _capi.SpectrogramParams._fields_]
    assert C.sizeof(_capi.SpectrogramParams) == 4 * 8 + 8 * 4 + 4 * 2 + 8 * 3


def _header_prototypes():
    """name -> (return layout, [argument layouts]) in the vocabulary of java.lang.foreign."""
    src = open(os.path.join(ROOT, "include", "sa_engine.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)

    def layout(t):
        t = t.strip()
        if "*" in t:
            return "ADDRESS"
        base = t.replace("const", "").split()[0]
        return {"int32_t": "JAVA_INT", "uint32_t": "JAVA_INT", "uint64_t": "JAVA_LONG", "double": "JAVA_DOUBLE",
                "void": "VOID"}[base]

    out = {}
    for ret, name, args in re.findall(r"SA_API\s+([\w\s\*]+?)\b(sa_\w+)\s*\((.*?)\)\s*;", src, re.S):
        args = [a for a in (x.strip() for x in args.split(",")) if a and a != "void"]
        # an argument is "type name": the layout comes from everything before the last identifier
        out[name] = (layout(ret), [layout(re.sub(r"\w+$", "", a) if not a.endswith("*") else a) for a in args])
    return out


def test_java_function_descriptors_match_the_header_prototypes():
    java = open(os.path.join(ROOT, "java", "net", "kcundercover", "spectral_analyzer", "services",
                             "NativeSpectralEngine.java")).read()
    protos = _header_prototypes()
    found = re.findall(r'fn\("(sa_\w+)",\s*FunctionDescriptor\.(ofVoid|of)\((.*?)\)\);', java, re.S)
    assert len(found) >= 12
    for name, kind, body in found:
        lay = [x.strip() for x in body.replace("\n", " ").split(",") if x.strip()]
        ret, args = ("VOID", lay) if kind == "ofVoid" else (lay[0], lay[1:])
        assert (ret, args) == protos[name], (name, ret, args, protos[name])


def test_ctypes_argtypes_match_the_header_prototypes():
    """Same check for the Python mirror: widths and pointer-ness of every argument and of the return type."""
    L = _capi.lib()

    def layout(ct):
        if ct is None:
            return "VOID"
        if ct in (C.c_int32, C.c_uint32, C.c_int, C.c_uint):
            return "JAVA_INT"
        if ct in (C.c_uint64, C.c_int64, C.c_size_t, C.c_ulonglong, C.c_longlong):
            return "JAVA_LONG"
        if ct is C.c_double:
            return "JAVA_DOUBLE"
        return "ADDRESS"                                   # c_void_p, c_char_p, POINTER(...), structure pointers

    for name, (ret, args) in _header_prototypes().items():
        f = getattr(L, name)
        if f.argtypes is None:
            assert not args, name
            continue
        assert [layout(a) for a in f.argtypes] == args, (name, f.argtypes, args)
        assert layout(f.restype) == ret, (name, f.restype, ret)


def test_cpp_host_mirror_compiles_and_links_on_cpu(tmp_path):
    """include/sa_services.hpp + tests/cpp/test_services.cpp build against libsa_engine.so with plain g++ (the binary is
    run by the GPU test; here only the build is checked, and that the header needs nothing from CUDA)."""
    import shutil
    import subprocess
    if not shutil.which("g++"):
        pytest.skip("no g++")
    exe = str(tmp_path / "test_services")
    lib_dir = os.path.join(ROOT, "spectral_analyzer_b200")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "cpp", "test_services.cpp"), "-o", exe, "-L" + lib_dir, "-lsa_engine",
                        "-Wl,-rpath," + lib_dir], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "warning" not in r.stderr, r.stderr[-2000:]
