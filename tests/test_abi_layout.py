"""CPU: ctypes mirrors of the C ABI's structs have the C compiler's layout (a drifted mirror reads garbage)."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_sizeof(tmp_path, header, names):
    src = tmp_path / "sz.c"
    body = "".join('printf("%%zu %%zu\\n", sizeof(%s), offsetof(%s, %s));\n' % (n, n, last) for n, last in names)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(void){%s return 0;}\n' % (header, body))
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    return [tuple(int(x) for x in line.split()) for line in out if line.strip()]


def test_aligner_info_struct_layout(tmp_path):
    from racon_gpu_b200.aligner import AlnBatchInfo
    (size, off), = _c_sizeof(tmp_path, "b200aln.h", [("b200aln_batch_info", "kernel_ms")])
    assert C.sizeof(AlnBatchInfo) == size and AlnBatchInfo.kernel_ms.offset == off


def test_poa_struct_layouts(tmp_path):
    from racon_gpu_b200 import api
    got = _c_sizeof(tmp_path, "b200poa.h", [("b200poa_config", "band_mode"), ("b200poa_entry", "end")])
    assert (C.sizeof(api.Config), api.Config.band_mode.offset) == got[0]
    assert (C.sizeof(api.Entry), api.Entry.end.offset) == got[1]
