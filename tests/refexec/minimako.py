"""A renderer for the subset of the Mako template language the reference's kernel templates use
(``mako`` is not installed in this image): ``${expr}``, ``<% python %>`` blocks,
``<%def name="f(args)"> ... </%def>`` (called as ``${f(...)}``), ``% if/elif/else/for/endif/endfor``
control lines, ``##`` comment lines, ``%%`` and backslash-newline joining.

Test infrastructure only: it exists so that ``tests/refexec`` can render the kernel sources that lie
under ``/root/reference`` unmodified and execute them on the CPU.  A template is translated to a
Python module once; rendering executes it with the render variables as globals, so undefined
names raise ``NameError`` (the reference always asks for ``strict_undefined=True``).
"""
from __future__ import annotations

import re
import textwrap

_CONTROL = re.compile(r"[ \t]*%(?!%)[ \t]*((?:[^\n\\]|\\[^\n]|\\\n)*)(?:\n|\Z)")
_COMMENT = re.compile(r"[ \t]*##[^\n]*(?:\n|\Z)")
_DEF_OPEN = re.compile(r"<%def\s+name\s*=\s*\"([^\"]*)\"\s*>", re.S)
_OPENERS = ("if", "for", "while", "try", "with")
_MIDDLES = ("elif", "else", "except", "finally")


def _scan_expression(src: str, pos: int) -> int:
    """Index of the ``}`` closing a ``${`` opened just before ``pos``."""
    depth = 0
    i = pos
    n = len(src)
    while i < n:
        c = src[i]
        if c in "\"'":
            quote = c
            if src.startswith(quote * 3, i):
                end = src.index(quote * 3, i + 3)
                i = end + 3
                continue
            i += 1
            while src[i] != quote:
                i += 2 if src[i] == "\\" else 1
            i += 1
            continue
        if c in "{[(":
            depth += 1
        elif c in ")]":
            depth -= 1
        elif c == "}":
            if depth == 0:
                return i
            depth -= 1
        i += 1
    raise SyntaxError("unterminated ${ in template")


class Template:
    """Mirrors the two things the reference uses: ``Template(text, strict_undefined=True)`` and
    ``.render(**vars)``."""

    def __init__(self, text: str, strict_undefined: bool = True, **_ignored):
        self.source = text
        self._code = compile(self._translate(text), "<minimako>", "exec")

    # {{{ translation

    def _translate(self, src: str) -> str:
        lines: list[str] = []
        indent = 0
        # a stack of indentation levels at which a <%def> body started
        def_stack: list[int] = []

        def emit(line: str) -> None:
            lines.append("    " * indent + line)

        def emit_text(text: str) -> None:
            if text:
                emit(f"__w({text!r})")

        pos = 0
        n = len(src)
        bol = True
        buf: list[str] = []

        def flush() -> None:
            emit_text("".join(buf))
            buf.clear()

        while pos < n:
            if bol:
                m = _COMMENT.match(src, pos)
                if m:
                    pos = m.end()
                    continue
                m = _CONTROL.match(src, pos)
                if m:
                    flush()
                    stmt = m.group(1).replace("\\\n", " ").strip()
                    pos = m.end()
                    keyword = re.match(r"[A-Za-z_]+", stmt)
                    kw = keyword.group(0) if keyword else ""
                    if kw.startswith("end") and kw[3:] in _OPENERS:
                        indent -= 1
                    elif kw in _MIDDLES:
                        indent -= 1
                        emit(stmt if stmt.endswith(":") else stmt + ":")
                        indent += 1
                        emit("pass")
                    elif kw in _OPENERS:
                        emit(stmt if stmt.endswith(":") else stmt + ":")
                        indent += 1
                        emit("pass")
                    else:
                        raise SyntaxError(f"unsupported control line: % {stmt}")
                    continue
                bol = False
            if src.startswith("${", pos):
                flush()
                end = _scan_expression(src, pos + 2)
                expr = src[pos + 2:end].strip()
                if "|" in expr and re.search(r"\|\s*[a-z, ]+$", expr):
                    expr = expr[:expr.rindex("|")].strip()
                emit(f"__w(__s({' '.join(expr.splitlines())}))")
                pos = end + 1
                continue
            m = _DEF_OPEN.match(src, pos)
            if m:
                flush()
                sig = " ".join(m.group(1).split())
                emit(f"def {sig}:")
                def_stack.append(indent)
                indent += 1
                emit("pass")
                pos = m.end()
                continue
            if src.startswith("</%def>", pos):
                flush()
                emit("return ''")
                indent = def_stack.pop()
                pos += len("</%def>")
                continue
            if src.startswith("<%", pos):
                flush()
                end = src.index("%>", pos)
                block = textwrap.dedent(src[pos + 2:end].replace("\\\n", " "))
                # a one-line block may start with a space: dedent handles whole blocks only
                block = "\n".join(ln for ln in block.splitlines() if ln.strip())
                block = textwrap.dedent(block)
                for ln in block.splitlines():
                    emit(ln)
                pos = end + 2
                continue
            if src.startswith("\\\n", pos):
                pos += 2                      # Mako joins the lines
                continue
            if src.startswith("%%", pos) and (pos == 0 or src[pos - 1] == "\n"):
                buf.append("%")
                pos += 2
                continue
            c = src[pos]
            buf.append(c)
            pos += 1
            if c == "\n":
                bol = True
        flush()
        if indent != 0 or def_stack:
            raise SyntaxError("unbalanced control lines / defs in template")
        return "\n".join(lines) + "\n"

    # }}}

    def render(self, **variables) -> str:
        out: list[str] = []
        namespace = dict(variables)
        namespace["__w"] = out.append
        namespace["__s"] = lambda v: v if isinstance(v, str) else str(v)
        exec(self._code, namespace)
        return "".join(out)
