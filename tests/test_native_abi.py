"""CPU tests: the C-ABI library loads and exports every symbol include/rloa_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from robotic_manipulator_rloa_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'rloa_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(rloa_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 30
    handle = ctypes.CDLL(_native.LIB_PATH)
    for name in names:
        assert hasattr(handle, name), f'{name} declared in rloa_b200.h but not exported'
    assert sorted(_native.SIGNATURES) == names, 'ctypes signature table out of sync with the header'
    lib = _native.lib()
    assert lib.rloa_version() == 100
    assert lib.rloa_launch_count() == 0 or lib.rloa_launch_count() > 0


def test_struct_sizes_match_header_layout():
    # rloa_step_config: 2 x (int + int[32]) + 4 floats ; rloa_naf_hyper: 7 floats + 2 ints + 1 float
    assert ctypes.sizeof(_native.StepConfig) == 4 * (1 + 32 + 1 + 32 + 4)
    assert ctypes.sizeof(_native.NafHyper) == 4 * 10
    assert ctypes.sizeof(_native.NafParams) == 16 + 20 * 8
    assert ctypes.sizeof(_native.Replay) == 16 + 7 * 8


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_native, 'LIB_PATH', str(tmp_path / 'nope.so'))
    monkeypatch.setattr(_native, '_lib', None)
    try:
        _native.lib()
    except _native.NativeLibraryError as err:
        assert 'no CPU fallback' in str(err)
    else:
        raise AssertionError('expected NativeLibraryError')


def test_ctypes_structs_match_the_header_as_a_c_compiler_sees_it(tmp_path):
    """sizeof / offsetof of the ABI structs from gcc on include/rloa_b200.h against the ctypes mirrors in _native.py."""
    import subprocess
    src = tmp_path / 'probe.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rloa_b200.h"\n'
                   'int main(void) {\n'
                   '  printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(rloa_model_desc), offsetof(rloa_model_desc, base_R),\n'
                   '         offsetof(rloa_model_desc, ns), offsetof(rloa_model_desc, ee_link), offsetof(rloa_model_desc, n_verts),\n'
                   '         offsetof(rloa_model_desc, verts), sizeof(rloa_step_config), sizeof(rloa_naf_hyper));\n'
                   '  return 0;\n}\n')
    exe = tmp_path / 'probe'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    M = _native.ModelDesc
    want = [ctypes.sizeof(M), M.base_R.offset, M.ns.offset, M.ee_link.offset, M.n_verts.offset, M.verts.offset,
            ctypes.sizeof(_native.StepConfig), ctypes.sizeof(_native.NafHyper)]
    assert got == want


def test_binding_table_has_the_header_s_parameter_counts():
    """Every prototype in include/rloa_b200.h against the ctypes table: same number of parameters, and pointer /
    integer / float kinds in the same positions (a changed C signature must not go unnoticed on a machine without a GPU)."""
    text = open(os.path.join(ROOT, 'include', 'rloa_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    protos = re.findall(r'\b(rloa_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', text, flags=re.S)
    assert len(protos) >= 30
    for name, params in protos:
        params = ' '.join(params.split())
        plist = [] if params in ('', 'void') else [p.strip() for p in params.split(',')]
        restype, argtypes = _native.SIGNATURES[name]
        assert len(argtypes) == len(plist), f'{name}: header has {len(plist)} parameters, the binding {len(argtypes)}'
        for p, a in zip(plist, argtypes):
            is_ptr = '*' in p
            ptype = p.rsplit(' ', 1)[0]                      # drop the parameter name
            if is_ptr:
                assert a in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(a, 'contents') or hasattr(a, '_type_'), (name, p, a)
            elif 'float' in ptype or 'double' in ptype:
                assert a in (ctypes.c_float, ctypes.c_double), (name, p, a)
            else:
                assert a in (ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_size_t), (name, p, a)
