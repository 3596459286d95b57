"""Development-time helper: lists BSIM4 model parameter names and their default values as an X-macro table.

The 850-odd BSIM4.8 model parameters and their defaults are data (they come from the Berkeley BSIM4 manual; the
reference restates them in spice21/src/comps/bsim4/model/vals.rs:9-1376). Typing them by hand would only add typos, so
this script reads the reference's list ONCE, in declaration order, and writes a compact table that
spice21_b200/csrc/bsim4/bsim4_model.hpp expands into a struct and a default-resolution loop:

    B4P(name, literal)        default is a literal
    B4A(name, other)          default is the (already resolved) value of another parameter
    B4T(name, nmos, pmos)     default depends on the device polarity

Everything that is not a plain default (clamps, derived capacitances, the aigsd/bigsd/cigsd groups, given-flags) is
written by hand in bsim4_model.hpp. Run from the repo root; needs /root/reference.
"""
import re

SRC = "/root/reference/spice21/src/comps/bsim4/model/vals.rs"
OUT = "spice21_b200/csrc/bsim4/bsim4_model_table.inc"
HAND = {"tnom", "mos_type", "ua", "uc", "uc1", "cf", "cgso", "cgdo", "cgbo", "aigs", "aigd", "bigs", "bigd", "cigs", "cigd",
        "aigsd", "bigsd", "cigsd"}

src = open(SRC).read()
one = re.compile(r"vals\.(r#)?(\w+) = if let Some\(val\) = specs\.(?:r#)?\w+ \{ val(?: as usize)? \} else \{ ([^}]*) \};")
typed = re.compile(r"vals\.(\w+) = if let Some\(val\) = specs\.\w+ \{\s*val\s*\} else \{\s*match vals\.mos_type \{\s*NMOS => ([^,]+),\s*PMOS => ([^,]+),\s*\}\s*\};")
items = []
for m in one.finditer(src):
    items.append((m.start(), m.group(2), "one", m.group(3).strip()))
for m in typed.finditer(src):
    items.append((m.start(), m.group(1), "typed", (m.group(2).strip(), m.group(3).strip())))
items.sort()
lines, seen = [], set()
for _, name, kind, d in items:
    if name in HAND or name in seen:
        continue
    seen.add(name)
    if kind == "typed":
        lines.append(f"B4T({name}, {d[0]}, {d[1]})")
    elif d.startswith("vals."):
        lines.append(f"B4A({name}, {d[5:]})")
    else:
        float(d)
        lines.append(f"B4P({name}, {d})")
with open(OUT, "w") as f:
    f.write("// BSIM4 model parameters and defaults, in resolution order (see scripts/gen_bsim4_tables.py).\n")
    f.write("// B4P(name, default)  B4A(name, defaults-to-other-param)  B4T(name, nmos default, pmos default)\n")
    f.write("\n".join(lines) + "\n")
print(len(lines), "parameters ->", OUT)

# ---- second table: the size-binned parameters, p = base + l*Inv_L + w*Inv_W + p*Inv_LW (bsim4inst.rs:234-391), in order.
SRC2 = "/root/reference/spice21/src/comps/bsim4/bsim4inst.rs"
OUT2 = "spice21_b200/csrc/bsim4/bsim4_binned_table.inc"
pat = re.compile(r"size_params\.(\w+) = model\.(?:r#)?(\w+) \+ model\.(?:r#)?(\w+) \* Inv_L \+ model\.(?:r#)?(\w+) \* Inv_W \+ model\.(?:r#)?(\w+) \* Inv_LW;")
rows = pat.findall(open(SRC2).read())
with open(OUT2, "w") as f:
    f.write("// Size-binned BSIM4 parameters: B4BIN(size-dependent field, model base, l-term, w-term, p-term), in evaluation order.\n")
    for r in rows:
        f.write("B4BIN(%s, %s, %s, %s, %s)\n" % r)
print(len(rows), "binned parameters ->", OUT2)

# ---- third table: field inventories of the per-size and per-instance precomputed blocks (bsim4/mod.rs:69-164, 441-676),
# so the flat device parameter block (bsim4_params.hpp) can be laid out without typing ~350 names.
SRC3 = "/root/reference/spice21/src/comps/bsim4/mod.rs"
OUT3 = "spice21_b200/csrc/bsim4/bsim4_fields.inc"
mod = open(SRC3).read()
def fields(struct):
    body = mod[mod.index("struct " + struct):]
    body = body[:body.index("\n}")]
    return re.findall(r"pub\(crate\) (?:r#)?(\w+): (\w+)", body)
with open(OUT3, "w") as f:
    f.write("// Field inventories: B4S(name) size-dependent block, B4I(name) per-instance block (all stored as double).\n")
    for n, _ in fields("Bsim4SizeDepParams"):
        f.write("B4S(%s)\n" % n)
    for n, _ in fields("Bsim4InternalParams"):
        f.write("B4I(%s)\n" % n)
print("field inventories ->", OUT3)
