"""rust/motivate-b200-sys/src/lib.rs against include/motivate_b200.h, mechanically (there is no cargo/rustc in this
image, so nothing compiles the crate): same set of functions, same argument count and types in the same order, same
return types, same struct fields in the same order, same constants."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "motivate_b200.h")
LIB = os.path.join(ROOT, "rust", "motivate-b200-sys", "src", "lib.rs")

C2RUST = {"uint8_t": "u8", "uint16_t": "u16", "uint32_t": "u32", "uint64_t": "u64", "int": "c_int", "float": "f32",
          "double": "f64", "size_t": "usize", "char": "c_char", "void": "c_void"}


def c_type_to_rust(t):
    """'const uint8_t*' -> '*const u8', 'mvb_sim**' -> '*mut *mut mvb_sim', 'const mvb_population* const*' -> '*const *const mvb_population'."""
    t = t.strip()
    toks = re.findall(r"[A-Za-z_0-9]+|\*", t)
    base = [x for x in toks if x not in ("const", "*", "struct")]
    assert len(base) == 1, t
    base = C2RUST.get(base[0], base[0])
    # walk the declarator: each '*' is const if a 'const' sits right of it, else takes the const to the left of the base for the innermost
    ptrs = []
    seen_base_const = False
    i = 0
    while i < len(toks):
        if toks[i] == "const" and not ptrs and (i == 0 or toks[i - 1] != "*"):
            seen_base_const = True
        if toks[i] == "*":
            ptrs.append(i + 1 < len(toks) and toks[i + 1] == "const")
        i += 1
    out = base
    for k, p in enumerate(ptrs):
        pointee_const = seen_base_const if k == 0 else ptrs[k - 1]
        out = ("*const " if pointee_const else "*mut ") + out
    return out


def split_args(s):
    s = s.strip()
    if s in ("", "void"):
        return []
    return [a.strip() for a in s.split(",")]


def parse_header():
    src = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    structs = {}
    for body, name in re.findall(r"typedef struct \w+ \{(.*?)\}\s*(\w+);", src, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(.*?)(\w+)((?:\[\d+\])*)$", decl, flags=re.S)
            ctype, fname, dims = m.group(1), m.group(2), re.findall(r"\[(\d+)\]", m.group(3))
            rt = c_type_to_rust(ctype)
            for d in reversed(dims):
                rt = f"[{rt}; {d}]"
            fields.append((fname, rt))
        structs[name] = fields
    src_nostruct = re.sub(r"typedef struct \w+ \{.*?\}\s*\w+;", "", src, flags=re.S)
    funcs = {}
    for ret, name, args in re.findall(r"([A-Za-z_][A-Za-z_0-9 \*]*?)\s*\b(mvb_\w+)\s*\(([^)]*)\)\s*;", src_nostruct):
        a = []
        for arg in split_args(args):
            m = re.match(r"(.*?)(\w+)$", arg, flags=re.S)
            a.append(c_type_to_rust(m.group(1)))
        ret = ret.strip()
        funcs[name] = (a, None if ret == "void" else c_type_to_rust(ret))
    consts = {k: v for k, v in re.findall(r"^\s*#define\s+(MVB_\w+)\s+(-?\d+)u?\b", open(HDR).read(), flags=re.M)}
    return structs, funcs, consts


def parse_rust():
    src = re.sub(r"//.*$", "", open(LIB).read(), flags=re.M)
    raw = src[src.index("pub mod raw"):]
    structs = {}
    for name, body in re.findall(r"pub struct (mvb_\w+) \{(.*?)\n    \}", raw, flags=re.S):
        structs[name] = [(f, t.strip()) for f, t in re.findall(r"pub (\w+):\s*([^,\n]+),", body)]
    ext = raw[raw.index('extern "C"'):]
    ext = ext[:ext.index("\n    }")]
    funcs = {}
    for name, args, ret in re.findall(r"pub fn (mvb_\w+)\((.*?)\)\s*(?:->\s*([^;]+))?;", ext, flags=re.S):
        a = [x.split(":", 1)[1].strip() for x in split_args(args)]
        funcs[name] = (a, ret.strip() if ret else None)
    consts = {k: v for k, v in re.findall(r"pub const (MVB_\w+):\s*\w+\s*=\s*(-?\d+);", src)}
    return structs, funcs, consts


def test_type_mapper():
    assert c_type_to_rust("const uint8_t* ") == "*const u8"
    assert c_type_to_rust("mvb_sim** ") == "*mut *mut mvb_sim"
    assert c_type_to_rust("const mvb_population* const* ") == "*const *const mvb_population"
    assert c_type_to_rust("uint64_t** ") == "*mut *mut u64"
    assert c_type_to_rust("const void* ") == "*const c_void"


def test_rust_extern_block_matches_the_header():
    hs, hf, hc = parse_header()
    rs, rf, rc = parse_rust()
    assert len(hf) >= 31
    assert set(hf) == set(rf), f"only in header: {sorted(set(hf) - set(rf))}; only in lib.rs: {sorted(set(rf) - set(hf))}"
    for name, (args, ret) in hf.items():
        assert rf[name][0] == args, f"{name}: header {args} vs lib.rs {rf[name][0]}"
        assert rf[name][1] == ret, f"{name}: return {ret} vs {rf[name][1]}"


def test_rust_structs_match_the_header():
    hs, _, _ = parse_header()
    rs, _, _ = parse_rust()
    assert set(hs) == set(rs), (sorted(hs), sorted(rs))
    for name, fields in hs.items():
        assert rs[name] == fields, f"{name}: header {fields} vs lib.rs {rs[name]}"


def test_rust_constants_match_the_header():
    _, _, hc = parse_header()
    _, _, rc = parse_rust()
    for k in ("MVB_ABI_VERSION", "MVB_OK", "MVB_ERR_ARG", "MVB_ERR_CUDA", "MVB_ERR_NOMEM", "MVB_ERR_STATE", "MVB_ERR_DOMAIN",
              "MVB_ERR_COMM", "MVB_N_MODES", "MVB_N_COMMUTES", "MVB_STATS_FIXED", "MVB_UNIQUE_ID_BYTES", "MVB_FLAG_NONE",
              "MVB_EXCHANGE_NONE", "MVB_EXCHANGE_NCCL", "MVB_EXCHANGE_P2P"):
        assert hc[k] == rc[k], (k, hc.get(k), rc.get(k))


def test_patch_uses_only_wrapper_items_that_exist():
    """rust/patch/simulation_b200.rs calls the safe wrapper: every `mvb::X` / `sim.method(` it names is defined in lib.rs."""
    lib = open(LIB).read()
    patch = open(os.path.join(ROOT, "rust", "patch", "simulation_b200.rs")).read()
    for item in set(re.findall(r"mvb::(\w+)", patch)):
        assert re.search(rf"pub (struct|fn|mod|enum) {item}\b", lib), item
    for meth in set(re.findall(r"\bsim\.(\w+)\(", patch)):
        assert re.search(rf"pub fn {meth}\b", lib), meth
    for field in set(re.findall(r"\bp\.(\w+)\.push", patch)) | set(re.findall(r"\bt\.(\w+)\.push", patch)):
        assert re.search(rf"pub {field}:", lib), field
