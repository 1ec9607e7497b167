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
    # rloa_step_config: 2 x (int + int[32]) + 3 floats ; rloa_naf_hyper: 7 floats + 2 ints + 1 float
    assert ctypes.sizeof(_native.StepConfig) == 4 * (1 + 32 + 1 + 32 + 3)
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
