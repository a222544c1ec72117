"""CPU: the C-ABI library loads without a GPU, exports every symbol include/b200_expect.h declares,
cross-compiles models for sm_100a through NVRTC, validates its inputs, and REFUSES to compute without a
device (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200_expect.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200e_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(sme):
    L = sme.load_library()
    names = _declared_symbols()
    assert len(names) >= 21
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/b200_expect.h but not exported"
    assert sorted(sme._capi.EXPORTS) == names
    assert L.b200e_abi_version() == sme._capi.ABI_VERSION == 4


def test_struct_layout_matches_header(sme):
    # sizes computed by hand from the header (LP64): catches drift between the ctypes mirror and the C structs
    assert C.sizeof(sme._capi.PdfDim) == 56
    assert C.sizeof(sme._capi.Counters) == 32
    assert C.sizeof(sme._capi.CallStats) == 64
    assert C.sizeof(sme._capi.Model) == 176
    assert C.sizeof(sme._capi.CubatureOpts) == 24 and C.sizeof(sme._capi.CubatureStats) == 88


def test_field_offsets_match_a_c_compile_of_the_header(sme, tmp_path):
    """The ctypes mirror (and with it the Julia shim, which lists the same fields in the same order) against
    offsetof() of every field of every struct, from a C program that includes include/b200_expect.h."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    structs = {"b200e_pdf_dim": sme._capi.PdfDim, "b200e_model": sme._capi.Model, "b200e_counters": sme._capi.Counters,
               "b200e_call_stats": sme._capi.CallStats, "b200e_cubature_opts": sme._capi.CubatureOpts,
               "b200e_cubature_stats": sme._capi.CubatureStats}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200_expect.h"', "int main(void) {"]
    for cname, ct in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "offsets.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "offsets"
    subprocess.run([cc, "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, ct in structs.items():
        assert int(out[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(ct, fname).offset, (cname, fname)


def test_compile_check_all_builtin_models_for_sm100a(sme, tmp_path):
    for name in sme.models.BUILTIN:
        for fp in ((0, 1) if name in ("linear2", "lorenz63", "oscillator_noise") else (0,)):   # keep the CPU suite short
            spec = sme.models.builtin(name, fp_mode=fp, cache_dir=str(tmp_path))
            nbytes, log = sme.compile_check(spec)
            assert nbytes > 10000
            assert "sm_100a" in log and "b200e_reduce" in log and "b200e_nodes" in log
            spills = [int(v) for v in re.findall(r"(\d+) bytes spill stores", log)]
            # the library picks the densest residency (4 / 3 / 2 blocks of 256) within its 32-byte spill allowance;
            # a system too large for 128 registers (8 states in FP64) stays at two blocks and spills into L1
            allowance = 32 if spec.nstate * (1 if fp else 2) < 16 else 512
            if spec.solver == sme.models.SOLVER_LAMBA_EM:
                allowance = 64   # measured: three blocks with up to 52 B of spills beat two blocks without (DESIGN K8)
            assert len(spills) == 5 and max(spills) <= allowance, (name, fp, spills)   # nodes / reduce x static / dynamic + reduce_gen
            # second call is served from the cubin cache
            n2, log2 = sme.compile_check(spec)
            assert n2 == nbytes and "loaded from cache" in log2


def test_compile_error_is_reported_with_log(sme):
    bad = sme.models.builtin("linear2", cache_dir="")
    bad = bad.replace(source=bad.source.replace("du[0] = u[1];", "du[0] = undefined_symbol;"))
    with pytest.raises(sme.B200Error) as e:
        sme.compile_check(bad)
    assert e.value.code == sme._capi.ERR_COMPILE and "undefined_symbol" in str(e.value)


@pytest.mark.parametrize("field,value", [("nstate", 0), ("nstate", 17), ("nx", 33), ("nout", 0), ("solver", 7),
                                         ("fp_mode", 3), ("block_size", 48), ("tspan", (1.0, 1.0)),
                                         ("abstol", -1.0), ("refill_threshold", 40)])
def test_invalid_models_are_rejected(sme, field, value):
    spec = sme.models.builtin("linear2").replace(**{field: value})
    with pytest.raises(sme.B200Error) as e:
        sme.compile_check(spec)
    assert e.value.code == sme._capi.ERR_INVALID


def test_invalid_pdf_and_em_settings(sme):
    spec = sme.models.builtin("linear2")
    with pytest.raises(sme.B200Error):
        sme.compile_check(spec.replace(pdf=[(1, 3.0, 2.0, 0, 1), (3, 0.0, 6.0, 3.0, 1.0)]))
    with pytest.raises(sme.B200Error):
        sme.compile_check(spec.replace(pdf=[(1, 1.0, 10.0, 0, 1), (3, 0.0, 6.0, 3.0, -1.0)]))
    em = sme.models.builtin("gbm4")
    with pytest.raises(sme.B200Error):
        sme.compile_check(em.replace(dt_fixed=0.0))
    with pytest.raises(sme.B200Error):
        sme.compile_check(em.replace(n_kkl=0))


def test_no_cpu_fallback(sme):
    """On a machine without a CUDA device every compute entry point fails loudly."""
    if sme.device_count() > 0:
        pytest.skip("a CUDA device is visible here")
    with pytest.raises(sme.B200Error) as e:
        sme.Handle(sme.models.builtin("linear2"))
    assert e.value.code == sme._capi.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    with pytest.raises(sme.B200Error):
        sme.measure_fma_peak()
    with pytest.raises(sme.B200Error):
        sme.PinnedArray((4, 2))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "scimlexpectations.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "expect_oracle" not in text, f


def _julia_structs(text):
    """struct Name ... end blocks of julia/B200Expectations.jl -> {name: [(field, type), ...]} (one-line and
    multi-line forms, `a::T; b::T` separators, trailing comments)."""
    out = {}
    for m in re.finditer(r"^struct\s+(\w+)\s*(?:#[^\n]*)?\n(.*?)^end\b|^struct\s+(\w+);(.*?);\s*end", text, flags=re.S | re.M):
        name, body = (m.group(1), m.group(2)) if m.group(1) else (m.group(3), m.group(4))
        body = re.sub(r"#[^\n]*", "", body)
        out[name] = [(f, t) for f, t in re.findall(r"(\w+)::([\w{}]+)", body)]
    return out


def test_julia_shim_matches_the_header(sme):
    """julia/B200Expectations.jl cannot run in this image; what can be checked statically is checked: its struct
    definitions list the same fields, in the same order and with the same machine types, as the ctypes mirror (which
    `test_field_offsets_match_a_c_compile_of_the_header` ties to the C header), and every `ccall` names an exported
    entry point with as many arguments as the header declares."""
    text = open(os.path.join(ROOT, "julia", "B200Expectations.jl")).read()
    structs = _julia_structs(text)
    jl2c = {"Int32": C.c_int32, "UInt32": C.c_uint32, "Int64": C.c_int64, "UInt64": C.c_uint64, "Float64": C.c_double}
    pairs = {"PdfDim": sme._capi.PdfDim, "CModel": sme._capi.Model, "Counters": sme._capi.Counters,
             "CubatureOpts": sme._capi.CubatureOpts, "CubatureStats": sme._capi.CubatureStats}
    for jname, ct in pairs.items():
        assert jname in structs, (jname, sorted(structs))
        jf = structs[jname]
        assert [f for f, _ in jf] == [f for f, _ in ct._fields_], jname
        for (f, jt), (_, cty) in zip(jf, ct._fields_):
            if jt in jl2c:
                assert cty is jl2c[jt], (jname, f, jt, cty)
            elif jt == "Cstring" or jt.startswith("Ptr{"):
                assert C.sizeof(cty) == C.sizeof(C.c_void_p) and not issubclass(cty, C.Structure), (jname, f)
            else:
                assert jt == "Counters" and cty is sme._capi.Counters, (jname, f, jt)
    # every ccall: exported symbol, argument count = parameter count of the declaration in the header
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "b200_expect.h")).read(), flags=re.S)
    decl = {m.group(1): m.group(2) for m in re.finditer(r"\b(b200e_\w+)\s*\(([^;{]*?)\)\s*;", header)}
    calls = re.findall(r"ccall\(\(:(b200e_\w+), LIB\),\s*\w+,\s*\(([^)]*)\)", text)
    assert len(calls) >= 8
    for name, argt in calls:
        assert name in decl, f"{name} is not declared in include/b200_expect.h"
        n_jl = len([a for a in argt.split(",") if a.strip()])
        n_c = 0 if decl[name].strip() in ("", "void") else len(decl[name].split(","))
        assert n_jl == n_c, (name, argt, decl[name])


def test_julia_shim_names_are_imported_and_handles_are_released():
    """Static checks a Julia parser would make at load time, since the module cannot be loaded here: every
    qualified `Module.name` the shim uses refers to a module it imports (round 1 shipped `KernelDensity.UnivariateKDE`
    without importing KernelDensity: an UndefVarError on `using`), no method reaches into `pdf_func.dists` as a
    default argument, every handle is wrapped in the finalised `Handle` type (no raw b200e_create result is kept),
    and `quadalg` is taken as the keyword the reference's solve uses (src/expectation.jl:336-343)."""
    text = open(os.path.join(ROOT, "julia", "B200Expectations.jl")).read()
    code = "\n".join(line.split("#")[0] if not line.lstrip().startswith('"') else line for line in text.splitlines())
    code = re.sub(r'"(?:[^"\\]|\\.)*"', '""', code)  # string literals carry C source and messages
    imported = {"Base", "Core", "GC"}  # always in scope (GC is Base.GC, exported)
    for m in re.finditer(r"^\s*(?:using|import)\s+([^\n]+)", code, flags=re.M):
        spec = m.group(1).split(":")[0]
        imported.update(s.strip() for s in spec.split(",") if s.strip())
    used = set(re.findall(r"\b([A-Z]\w*)\.[A-Za-z_@]", code))
    # local bindings that hold structs (field access, not module access)
    local = {"S", "D", "L", "P"}
    missing = sorted(u for u in used - imported - local)
    assert not missing, f"qualified names from modules that are never imported: {missing}"
    assert "KernelDensity" in imported
    assert "pdf_func.dists" not in code, "reach for the density factors through density_factors(), not a default argument"
    # one b200e_create call site, and its result goes straight into a finalised Handle
    assert len(re.findall(r":b200e_create\b", code)) == 1
    assert re.search(r"finalizer\(release!, h\)", code) and "Handle(h[])" in code
    # b200e_destroy is called from release! only
    assert len(re.findall(r":b200e_destroy\b", code)) == 1
    # the reference's keyword, on both problem kinds
    assert len(re.findall(r"solve\(exprob::B200\w*Problem, expalg::Koopman, args\.\.\.; quadalg = HCubatureJL\(\)", code)) == 2
    # build_integrand keeps the reference's five positional arguments and no keywords (src/expectation.jl:169-172)
    for m in re.finditer(r"function build_integrand\(([^)]*)\)", code):
        assert ";" not in m.group(1) and len(m.group(1).split(",")) == 5, m.group(0)
