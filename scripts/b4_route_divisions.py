"""Rewrite every f64 division `X / Y` in the BSIM4 evaluation headers as `B4_DIV(X, Y)`.

Why: the device code can then route all ~400 divisions of one BSIM4 evaluation through csrc/scalar.h (exact split
division with the zero-numerator shortcut, or its branch-free deferred-exception form) by redefining ONE macro, instead
of the compiler's inline sequence + slow-path call per site. With the default definition `((a) / (b))` the generated
code is unchanged — and this tool must not change a single bit either, so it is strict about C++ grouping:

  * `/` and `*` are left-associative at one precedence level: the left operand of a `/` is the WHOLE multiplicative chain
    to its left (`a * b / c` -> `B4_DIV(a * b, c)`), the right operand is the next unary expression only
    (`a / b * c` -> `B4_DIV(a, b) * c`); divisions are rewritten leftmost first so chains nest (`a / b / c` ->
    `B4_DIV(B4_DIV(a, b), c)`).
  * a unary sign in front of the chain's first factor binds tighter than `*` and `/`, so it belongs to the operand.
  * preprocessor lines and comments are left alone.

`python scripts/b4_route_divisions.py --selftest` checks the rewriting on random expressions (original vs rewritten text
evaluated with identical operand values must agree bit for bit); `--apply` rewrites the headers in place.
"""
import os
import random
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["bsim4_eval.hpp", "bsim4_eval_channel.hpp", "bsim4_eval_leak.hpp", "bsim4_eval_charge.hpp", "bsim4_eval_stamp.hpp"]

TOKEN = re.compile(r"""
    (?P<ws>\s+)
  | (?P<lc>//[^\n]*)
  | (?P<bc>/\*.*?\*/)
  | (?P<pp>^[ \t]*\#(?:[^\n\\]|\\\n|\\.)*)
  | (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?[fFlLuU]*)
  | (?P<id>[A-Za-z_]\w*)
  | (?P<str>"(?:[^"\\]|\\.)*")
  | (?P<op>->|::|<<=|>>=|<<|>>|<=|>=|==|!=|&&|\|\||\+=|-=|\*=|/=|%=|&=|\|=|\^=|\+\+|--|[-+*/%<>=!~&|^?:;,.(){}\[\]])
""", re.X | re.S | re.M)


def tokenize(src):
    toks, pos = [], 0
    while pos < len(src):
        m = TOKEN.match(src, pos)
        if not m:
            raise ValueError(f"cannot tokenize at {pos}: {src[pos:pos + 40]!r}")
        toks.append((m.lastgroup, m.group()))
        pos = m.end()
    return toks


SKIP = ("ws", "lc", "bc")
OPEN, CLOSE = {"(": ")", "[": "]", "{": "}"}, {")": "(", "]": "[", "}": "{"}


def code_index(toks):
    return [i for i, (k, _) in enumerate(toks) if k not in SKIP and k != "pp"]


def match_forward(toks, ci, p):
    """ci[p] is an opening bracket: position (in ci) of its partner."""
    depth = 0
    for q in range(p, len(ci)):
        t = toks[ci[q]][1]
        if t in OPEN:
            depth += 1
        elif t in CLOSE:
            depth -= 1
            if depth == 0:
                return q
    raise ValueError("unbalanced")


def match_backward(toks, ci, p):
    depth = 0
    for q in range(p, -1, -1):
        t = toks[ci[q]][1]
        if t in CLOSE:
            depth += 1
        elif t in OPEN:
            depth -= 1
            if depth == 0:
                return q
    raise ValueError("unbalanced")


def primary_forward(toks, ci, p):
    """Unary expression starting at ci[p]: returns last position of it."""
    while toks[ci[p]][1] in ("-", "+", "!", "*", "&"):  # unary sign / not / dereference / address-of
        p += 1
    kind, t = toks[ci[p]]
    if t == "(":
        p = match_forward(toks, ci, p)
    elif kind in ("num", "id"):
        pass
    else:
        raise ValueError(f"unexpected right operand {t!r}")
    while p + 1 < len(ci):  # postfix: call, index, member
        nt = toks[ci[p + 1]][1]
        if nt in ("(", "["):
            if toks[ci[p]][0] == "num":
                break
            p = match_forward(toks, ci, p + 1)
        elif nt in (".", "->", "::") and toks[ci[p + 2]][0] == "id":
            p += 2
        else:
            break
    return p


KEYWORDS = ("return", "if", "while", "for", "else", "switch")


def primary_backward(toks, ci, p):
    """Postfix/primary expression ending at ci[p]: returns first position of it (without unary prefix)."""
    while True:
        kind, t = toks[ci[p]]
        if t in (")", "]"):
            q = match_backward(toks, ci, p)
            if t == ")" and q > 0 and toks[ci[q - 1]][0] == "id" and toks[ci[q - 1]][1] not in KEYWORDS:
                p = q - 1  # call: f(...)
            elif t == "]":
                p = q - 1  # index: base[...]
                continue
            else:
                return q   # parenthesised expression
        elif kind not in ("num", "id"):
            raise ValueError(f"unexpected left operand end {t!r}")
        if p >= 2 and toks[ci[p - 1]][1] in (".", "->", "::"):
            p -= 2
            continue
        return p


STOP_BEFORE_UNARY = {"(", ",", "=", "?", ":", "return", "{", ";", "}", "<", ">", "<=", ">=", "==", "!=", "&&", "||", "+", "-", "*", "/", "+=", "-=", "*=", "!",
                     "[", "else"}


def left_operand(toks, ci, p):
    """Multiplicative chain ending at ci[p] (the token before the `/`): returns its first position. A sign in front of a
    factor is included when it is a unary sign (preceded by an operator or an opening); a binary + / - ends the chain."""
    while True:
        p = primary_backward(toks, ci, p)
        while p > 0 and toks[ci[p - 1]][1] in ("-", "+", "!", "*", "&") and (p - 1 == 0 or toks[ci[p - 2]][1] in STOP_BEFORE_UNARY):
            p -= 1  # unary sign / dereference of this factor
        if p > 0 and toks[ci[p - 1]][1] in ("*", "%"):
            p -= 2
            continue
        if p > 0 and toks[ci[p - 1]][1] == "/":
            raise AssertionError("divisions must be rewritten leftmost first")
        return p


def rewrite(src, macro="B4_DIV"):
    n = 0
    while True:
        toks = tokenize(src)
        ci = code_index(toks)
        hit = next((p for p in range(len(ci)) if toks[ci[p]] == ("op", "/")), None)
        if hit is None:
            return src, n
        lo = left_operand(toks, ci, hit - 1)
        hi = primary_forward(toks, ci, hit + 1)
        text = lambda a, b: "".join(t for _, t in toks[ci[a]:ci[b] + 1])
        new = f"{macro}({text(lo, hit - 1).strip()}, {text(hit + 1, hi).strip()})"
        src = "".join(t for _, t in toks[:ci[lo]]) + new + "".join(t for _, t in toks[ci[hi] + 1:])
        n += 1


# ------------------------------------------------------------------------------------------------ self-test
def _rand_expr(rng, depth, names):
    if depth <= 0 or rng.random() < 0.25:
        r = rng.random()
        if r < 0.6 or depth <= 0:
            return rng.choice(names)
        if r < 0.8:
            return rng.choice(["2.0", "0.5", "3.0", "1.0e-3"])
        return "f(" + _rand_expr(rng, depth - 1, names) + ")" if rng.random() < 0.5 else "s.m"
    r = rng.random()
    a, b = _rand_expr(rng, depth - 1, names), _rand_expr(rng, depth - 1, names)
    if r < 0.35:
        return f"{a} / {b}"
    if r < 0.6:
        return f"{a} * {b}"
    if r < 0.7:
        return f"{a} + {b}"
    if r < 0.8:
        return f"{a} - {b}"
    if r < 0.9:
        return f"({a})"
    return f"-{a}" if not a.startswith("-") else f"({a})"


def selftest(trials=3000, seed=7):
    rng = random.Random(seed)
    names = ["a", "b", "c", "d", "T0", "T1"]

    class S:  # member access operand
        m = 1.75

    for k in range(trials):
        e = _rand_expr(rng, 5, names)
        stmt = f"x = {e};"
        out, _ = rewrite(stmt)
        env = {n: rng.choice([-1, 1]) * rng.uniform(0.1, 10.0) for n in names}
        env.update(f=lambda v: v * 1.25 + 0.5, s=S, B4_DIV=lambda p, q: p / q)
        try:
            want = eval(e, {}, env)
        except ZeroDivisionError:
            continue
        got = eval(out[len("x = "):-1], {}, env)
        assert want == got or (want != want and got != got), (e, out)
        assert "/" not in out, out
    print(f"selftest: {trials} random expressions rewritten, values identical")


def main():
    if "--selftest" in sys.argv:
        selftest()
        return
    d = os.path.join(ROOT, "spice21_b200", "csrc", "bsim4")
    total = 0
    for f in FILES:
        p = os.path.join(d, f)
        src = open(p).read()
        out, n = rewrite(src)
        total += n
        print(f"{f}: {n} divisions")
        if "--apply" in sys.argv:
            open(p, "w").write(out)
    print("total", total, "(applied)" if "--apply" in sys.argv else "(dry run)")


if __name__ == "__main__":
    main()
