"""The C-ABI library loads without a GPU and exports every symbol include/lst_b200.h declares."""
import ctypes as C
import os
import re

import pytest

from laser_spot_track_b200 import native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lst_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(lst_[a-z0-9_]+)\s*\(", src))
    return sorted(n for n in names if n not in ("lst_compute_fn",))


def test_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(built_lib)
    declared = _declared()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/lst_b200.h but not exported"
    assert sorted(N.EXPORTS) == declared, "native.EXPORTS out of sync with the header"


def test_struct_sizes_match_header(built_lib):
    # sizes a C compiler gives the header's structs (x86-64 SysV): checked once with gcc, frozen here
    import subprocess, tempfile, textwrap
    prog = textwrap.dedent("""
        #include <stdio.h>
        #include "lst_b200.h"
        int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(lst_params), sizeof(lst_template_match),
            sizeof(lst_record), sizeof(lst_reference_info), sizeof(lst_stats), sizeof(lst_task), sizeof(lst_task_result));return 0;}
    """)
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")], check=True)
        out = subprocess.run([os.path.join(d, "s")], check=True, capture_output=True, text=True).stdout.split()
    got = [C.sizeof(x) for x in (N.Params, N.TemplateMatch, N.Record, N.ReferenceInfo, N.Stats, N.Task, N.TaskResult)]
    assert got == [int(v) for v in out]


def test_version_and_no_gpu_is_loud(built_lib):
    lib = N.load()
    assert lib.lst_abi_version() == N.ABI_VERSION == 3
    if lib.lst_device_count() == 0:
        p = N.default_params(640, 480, N.PIX_U8)
        ctx = C.c_void_p()
        rc = lib.lst_create(C.byref(p), 0, C.byref(ctx))
        assert rc == N.LST_ERR_CUDA and not ctx.value
        assert b"no CPU fallback" in lib.lst_last_error(None)
        from laser_spot_track_b200 import LaserSpotTrack4, LstError
        with pytest.raises(LstError):
            LaserSpotTrack4(640, 480, 8)


def test_param_validation_mirrors_reference(built_lib):
    lib = N.load()
    ctx = C.c_void_p()
    for kw, code in [(dict(method=6), N.LST_ERR_INVALID),
                     (dict(method=-1), N.LST_ERR_INVALID),
                     (dict(pixel_type=12), N.LST_ERR_UNSUPPORTED),
                     (dict(pixel_type=N.PIX_U16, match_intensity=0), N.LST_ERR_UNSUPPORTED),
                     (dict(templ_size=4), N.LST_ERR_UNSUPPORTED),
                     (dict(width=0), N.LST_ERR_INVALID),
                     (dict(struct_size=8), N.LST_ERR_INVALID)]:
        p = N.default_params(640, 480, N.PIX_U8)
        for k, v in kw.items():
            setattr(p, k, v)
        assert lib.lst_create(C.byref(p), 0, C.byref(ctx)) == code, kw


def test_results_rows_match_the_python_table(built_lib):
    """one call fills the ResultsTable columns of LST4:864-875 for a batch; skipped
    frames add no row and no time step (LST4:845)."""
    from laser_spot_track_b200 import results_rows, results_table
    recs = (N.Record * 6)()
    for i, r in enumerate(recs):
        r.status = N.STATUS_SKIP if i in (1, 4) else N.STATUS_OK
        r.dX_pix, r.dY_pix, r.x_abs, r.y_abs, r.dL = 0.5 * i, -0.25 * i, 10.0 + i, 20.0 - i, 0.125 * i
    rows, src = results_rows(recs, time0=0.0, time_step=2.5, first_row_number=2)
    assert src.tolist() == [0, 2, 3, 5]
    want = results_table(recs, None, time_step=2.5)
    assert len(rows) == len(want) == 4
    for k, (row, w) in enumerate(zip(rows, want)):
        assert row[0] == w["Time"] and row[1] == k + 2
        assert (row[2], row[3], row[4], row[5], row[6]) == (w["dX_pix"], w["dY_pix"], w["X_abs"], w["Y_abs"], w["dL"])
