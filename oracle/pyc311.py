"""ORACLE tooling (test infrastructure): run the reference's OWN bytecode.

The reference's pure-PyTorch Mamba exists only as a CPython-3.11 bytecode file,
`/root/reference/models/mamba/__pycache__/simple_mamba.cpython-311.pyc`; this container has
Python 3.12 only, whose `marshal`/`co_code` mangles 3.11 code objects.  This module
therefore (1) parses the 3.11 marshal stream by hand into plain `Code` records and (2) interprets the subset of
3.11 opcodes those functions use, executing them against real torch / einops objects.  With it
`tests/golden/make_golden.py` produces fixtures that are outputs of the reference itself, and
`tests/test_oracle.py` pins the restatement in `oracle/simple_mamba.py` to them.

Nothing here is product code and nothing here copies reference source: it is an interpreter.
"""
from __future__ import annotations

import builtins
import operator
import struct
from dataclasses import dataclass, field
from typing import Any

# ----------------------------------------------------------------------------------------------------
# marshal reader (format version 4, CPython 3.11 code-object field order)
# ----------------------------------------------------------------------------------------------------
FLAG_REF = 0x80


@dataclass
class Code:
    argcount: int
    posonlyargcount: int
    kwonlyargcount: int
    stacksize: int
    flags: int
    code: bytes
    consts: tuple
    names: tuple
    localsplusnames: tuple
    localspluskinds: bytes
    filename: str
    name: str
    qualname: str
    firstlineno: int
    linetable: bytes
    exceptiontable: bytes
    extra: dict = field(default_factory=dict)


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.p, self.refs = data, 0, []

    def u8(self):
        v = self.d[self.p]
        self.p += 1
        return v

    def i32(self):
        v = struct.unpack_from("<i", self.d, self.p)[0]
        self.p += 4
        return v

    def raw(self, n):
        v = self.d[self.p:self.p + n]
        self.p += n
        return v

    def obj(self):
        b = self.u8()
        t, ref = chr(b & ~FLAG_REF), bool(b & FLAG_REF)
        idx = None
        if ref:  # reserve the slot first: containers may be referenced by their own children
            idx = len(self.refs)
            self.refs.append(None)
        v = self._read(t)
        if ref:
            self.refs[idx] = v
        return v

    def _read(self, t):
        if t == "0":
            return None
        if t == "N":
            return None
        if t == "T":
            return True
        if t == "F":
            return False
        if t == ".":
            return Ellipsis
        if t == "i":
            return self.i32()
        if t == "l":
            n = self.i32()
            digits = [struct.unpack_from("<H", self.raw(2))[0] for _ in range(abs(n))]
            v = sum(d << (15 * i) for i, d in enumerate(digits))
            return -v if n < 0 else v
        if t == "g":
            return struct.unpack("<d", self.raw(8))[0]
        if t == "y":
            return complex(*struct.unpack("<dd", self.raw(16)))
        if t == "s":
            return bytes(self.raw(self.i32()))
        if t in "tu":
            return self.raw(self.i32()).decode("utf-8", "surrogatepass")
        if t in "aA":
            return self.raw(self.i32()).decode("latin-1")
        if t in "zZ":
            return self.raw(self.u8()).decode("latin-1")
        if t == ")":
            return tuple(self.obj() for _ in range(self.u8()))
        if t == "(":
            return tuple(self.obj() for _ in range(self.i32()))
        if t == "[":
            return [self.obj() for _ in range(self.i32())]
        if t in "<>":
            items = [self.obj() for _ in range(self.i32())]
            return set(items) if t == "<" else frozenset(items)
        if t == "{":
            out = {}
            while True:
                k = self.obj()
                if k is None and self.d[self.p - 1] == ord("0"):
                    break
                out[k] = self.obj()
            return out
        if t == "r":
            return self.refs[self.i32()]
        if t == "c":
            ints = [self.i32() for _ in range(5)]
            code = self.obj()
            consts, names, lpn, lpk, fn, nm, qn = (self.obj() for _ in range(7))
            first = self.i32()
            lt, et = self.obj(), self.obj()
            return Code(*ints, code, consts, names, lpn, lpk, fn, nm, qn, first, lt, et)
        raise ValueError(f"marshal type {t!r} at {self.p - 1}")


def load_pyc(path) -> Code:
    data = open(path, "rb").read()
    magic = struct.unpack_from("<H", data, 0)[0]
    if not 3495 <= magic <= 3499:  # CPython 3.11a7 .. 3.11 final
        raise ValueError(f"{path}: magic {magic} is not a CPython 3.11 pyc")
    return _Reader(data[16:]).obj()


def find_code(root: Code, qualname: str) -> Code:
    """Depth-first search for a nested code object by qualified name (e.g. 'MambaBlock.selective_scan')."""
    stack = [root]
    while stack:
        c = stack.pop()
        if c.qualname == qualname:
            return c
        stack.extend(k for k in c.consts if isinstance(k, Code))
    raise KeyError(qualname)


# ----------------------------------------------------------------------------------------------------
# CPython 3.11 opcode numbers (Lib/opcode.py @ 3.11) — only what the interpreter implements
# ----------------------------------------------------------------------------------------------------
OP = {
    0: "CACHE", 1: "POP_TOP", 2: "PUSH_NULL", 9: "NOP", 10: "UNARY_POSITIVE", 11: "UNARY_NEGATIVE", 12: "UNARY_NOT",
    15: "UNARY_INVERT", 25: "BINARY_SUBSCR", 30: "GET_LEN", 60: "STORE_SUBSCR", 68: "GET_ITER", 83: "RETURN_VALUE",
    90: "STORE_NAME", 92: "UNPACK_SEQUENCE", 93: "FOR_ITER", 95: "STORE_ATTR", 97: "STORE_GLOBAL", 99: "SWAP",
    100: "LOAD_CONST", 101: "LOAD_NAME", 102: "BUILD_TUPLE", 103: "BUILD_LIST", 104: "BUILD_SET", 105: "BUILD_MAP",
    106: "LOAD_ATTR", 107: "COMPARE_OP", 110: "JUMP_FORWARD", 111: "JUMP_IF_FALSE_OR_POP", 112: "JUMP_IF_TRUE_OR_POP",
    114: "POP_JUMP_FORWARD_IF_FALSE", 115: "POP_JUMP_FORWARD_IF_TRUE", 116: "LOAD_GLOBAL", 117: "IS_OP",
    118: "CONTAINS_OP", 120: "COPY", 122: "BINARY_OP", 124: "LOAD_FAST", 125: "STORE_FAST", 126: "DELETE_FAST",
    128: "POP_JUMP_FORWARD_IF_NOT_NONE", 129: "POP_JUMP_FORWARD_IF_NONE", 133: "BUILD_SLICE", 135: "MAKE_CELL",
    136: "LOAD_CLOSURE", 137: "LOAD_DEREF", 138: "STORE_DEREF", 140: "JUMP_BACKWARD", 144: "EXTENDED_ARG",
    145: "LIST_APPEND", 149: "COPY_FREE_VARS", 151: "RESUME", 155: "FORMAT_VALUE", 156: "BUILD_CONST_KEY_MAP",
    157: "BUILD_STRING", 160: "LOAD_METHOD", 162: "LIST_EXTEND", 166: "PRECALL", 171: "CALL", 172: "KW_NAMES",
    173: "POP_JUMP_BACKWARD_IF_NOT_NONE", 174: "POP_JUMP_BACKWARD_IF_NONE", 175: "POP_JUMP_BACKWARD_IF_FALSE",
    176: "POP_JUMP_BACKWARD_IF_TRUE",
}
BINARY = {
    0: operator.add, 1: operator.and_, 2: operator.floordiv, 3: operator.lshift, 4: operator.matmul, 5: operator.mul,
    6: operator.mod, 7: operator.or_, 8: operator.pow, 9: operator.rshift, 10: operator.sub, 11: operator.truediv,
    12: operator.xor, 13: operator.iadd, 14: operator.iand, 15: operator.ifloordiv, 16: operator.ilshift,
    17: operator.imatmul, 18: operator.imul, 19: operator.imod, 20: operator.ior, 21: operator.ipow,
    22: operator.irshift, 23: operator.isub, 24: operator.itruediv, 25: operator.ixor,
}
COMPARE = {0: operator.lt, 1: operator.le, 2: operator.eq, 3: operator.ne, 4: operator.gt, 5: operator.ge}
_NULL = object()


class Function:
    """A reference code object bound to a globals dict; calling it interprets the 3.11 bytecode."""

    def __init__(self, code: Code, globs: dict, defaults=(), name=None):
        self.code, self.globs, self.defaults = code, globs, tuple(defaults)
        self.__name__ = name or code.name

    def __get__(self, obj, objtype=None):  # behaves as a method when stored on a class
        if obj is None:
            return self
        return lambda *a, **k: self(obj, *a, **k)

    def __call__(self, *args, **kwargs):
        return run(self.code, self.globs, args, kwargs, self.defaults)


def run(code: Code, globs: dict, args=(), kwargs=None, defaults=()):
    kwargs = kwargs or {}
    names_fast = code.localsplusnames
    nargs = code.argcount
    if code.kwonlyargcount or (code.flags & 0x0C):
        raise NotImplementedError(f"{code.qualname}: *args/**kwargs/kw-only parameters")
    fast: dict[str, Any] = {}
    if len(args) > nargs:
        raise TypeError(f"{code.qualname}() takes {nargs} positional arguments but {len(args)} were given")
    for i, a in enumerate(args):
        fast[names_fast[i]] = a
    for k, v in kwargs.items():
        if k not in names_fast[:nargs] or k in fast:
            raise TypeError(f"{code.qualname}() got an unexpected keyword argument {k!r}")
        fast[k] = v
    for i, dv in enumerate(defaults):
        fast.setdefault(names_fast[nargs - len(defaults) + i], dv)
    missing = [n for n in names_fast[:nargs] if n not in fast]
    if missing:
        raise TypeError(f"{code.qualname}() missing arguments {missing}")

    bc = code.code
    stack: list = []
    pc = 0  # in 2-byte code units
    kw_names = ()
    ext = 0
    n_units = len(bc) // 2
    while pc < n_units:
        op, arg = bc[2 * pc], bc[2 * pc + 1] | ext
        ext = 0
        pc += 1
        name = OP.get(op)
        if name is None:
            raise NotImplementedError(f"{code.qualname}: opcode {op} at unit {pc - 1}")
        if name in ("CACHE", "NOP", "RESUME", "PRECALL", "MAKE_CELL", "COPY_FREE_VARS"):
            continue
        if name == "EXTENDED_ARG":
            ext = arg << 8
        elif name == "POP_TOP":
            stack.pop()
        elif name == "PUSH_NULL":
            stack.append(_NULL)
        elif name == "LOAD_CONST":
            stack.append(code.consts[arg])
        elif name == "LOAD_FAST":
            stack.append(fast[names_fast[arg]])
        elif name == "STORE_FAST":
            fast[names_fast[arg]] = stack.pop()
        elif name == "DELETE_FAST":
            del fast[names_fast[arg]]
        elif name == "LOAD_GLOBAL":
            if arg & 1:
                stack.append(_NULL)
            nm = code.names[arg >> 1]
            stack.append(globs[nm] if nm in globs else getattr(builtins, nm))
        elif name == "LOAD_ATTR":
            stack.append(getattr(stack.pop(), code.names[arg]))
        elif name == "STORE_ATTR":
            obj = stack.pop()
            setattr(obj, code.names[arg], stack.pop())
        elif name == "LOAD_METHOD":
            obj = stack.pop()
            stack.append(_NULL)
            stack.append(getattr(obj, code.names[arg]))  # bound method: NULL + callable form
        elif name == "KW_NAMES":
            kw_names = code.consts[arg]
        elif name == "CALL":
            argv = [stack.pop() for _ in range(arg)][::-1]
            a1, a0 = stack.pop(), stack.pop()
            if a0 is _NULL:
                fn = a1
            else:  # (callable, self) form
                fn, argv = a0, [a1] + argv
            nkw = len(kw_names)
            kw = dict(zip(kw_names, argv[len(argv) - nkw:])) if nkw else {}
            pos = argv[:len(argv) - nkw] if nkw else argv
            kw_names = ()
            stack.append(fn(*pos, **kw))
        elif name == "BINARY_OP":
            b = stack.pop()
            a = stack.pop()
            stack.append(BINARY[arg](a, b))
        elif name == "BINARY_SUBSCR":
            k = stack.pop()
            stack.append(stack.pop()[k])
        elif name == "STORE_SUBSCR":
            k = stack.pop()
            obj = stack.pop()
            obj[k] = stack.pop()
        elif name == "UNARY_NEGATIVE":
            stack.append(-stack.pop())
        elif name == "UNARY_POSITIVE":
            stack.append(+stack.pop())
        elif name == "UNARY_NOT":
            stack.append(not stack.pop())
        elif name == "UNARY_INVERT":
            stack.append(~stack.pop())
        elif name == "COMPARE_OP":
            b = stack.pop()
            a = stack.pop()
            stack.append(COMPARE[arg](a, b))
        elif name == "IS_OP":
            b = stack.pop()
            a = stack.pop()
            stack.append((a is not b) if arg else (a is b))
        elif name == "CONTAINS_OP":
            b = stack.pop()
            a = stack.pop()
            stack.append((a not in b) if arg else (a in b))
        elif name == "BUILD_TUPLE":
            items = [stack.pop() for _ in range(arg)][::-1]
            stack.append(tuple(items))
        elif name == "BUILD_LIST":
            items = [stack.pop() for _ in range(arg)][::-1]
            stack.append(items)
        elif name == "BUILD_MAP":
            items = [stack.pop() for _ in range(2 * arg)][::-1]
            stack.append(dict(zip(items[0::2], items[1::2])))
        elif name == "BUILD_CONST_KEY_MAP":
            keys = stack.pop()
            vals = [stack.pop() for _ in range(arg)][::-1]
            stack.append(dict(zip(keys, vals)))
        elif name == "BUILD_SLICE":
            items = [stack.pop() for _ in range(arg)][::-1]
            stack.append(slice(*items))
        elif name == "BUILD_STRING":
            items = [stack.pop() for _ in range(arg)][::-1]
            stack.append("".join(items))
        elif name == "FORMAT_VALUE":
            spec = stack.pop() if arg & 4 else ""
            v = stack.pop()
            conv = arg & 3
            v = str(v) if conv == 1 else repr(v) if conv == 2 else ascii(v) if conv == 3 else v
            stack.append(format(v, spec))
        elif name == "LIST_APPEND":
            v = stack.pop()
            stack[-arg].append(v)
        elif name == "LIST_EXTEND":
            v = stack.pop()
            stack[-arg].extend(v)
        elif name == "UNPACK_SEQUENCE":
            items = list(stack.pop())
            if len(items) != arg:
                raise ValueError(f"unpack: expected {arg} values, got {len(items)}")
            stack.extend(items[::-1])
        elif name == "GET_ITER":
            stack.append(iter(stack.pop()))
        elif name == "FOR_ITER":
            try:
                stack.append(next(stack[-1]))
            except StopIteration:
                stack.pop()
                pc += arg
        elif name == "JUMP_FORWARD":
            pc += arg
        elif name == "JUMP_BACKWARD":
            pc -= arg
        elif name in ("POP_JUMP_FORWARD_IF_FALSE", "POP_JUMP_FORWARD_IF_TRUE", "POP_JUMP_BACKWARD_IF_FALSE",
                      "POP_JUMP_BACKWARD_IF_TRUE"):
            v = bool(stack.pop())
            if v == name.endswith("TRUE"):
                pc += arg if "FORWARD" in name else -arg
        elif name in ("POP_JUMP_FORWARD_IF_NONE", "POP_JUMP_FORWARD_IF_NOT_NONE", "POP_JUMP_BACKWARD_IF_NONE",
                      "POP_JUMP_BACKWARD_IF_NOT_NONE"):
            v = stack.pop() is None
            if v != name.endswith("NOT_NONE"):
                pc += arg if "FORWARD" in name else -arg
        elif name == "JUMP_IF_FALSE_OR_POP":
            if not stack[-1]:
                pc += arg
            else:
                stack.pop()
        elif name == "JUMP_IF_TRUE_OR_POP":
            if stack[-1]:
                pc += arg
            else:
                stack.pop()
        elif name == "COPY":
            stack.append(stack[-arg])
        elif name == "SWAP":
            stack[-1], stack[-arg] = stack[-arg], stack[-1]
        elif name == "GET_LEN":
            stack.append(len(stack[-1]))
        elif name == "RETURN_VALUE":
            return stack.pop()
        else:
            raise NotImplementedError(f"{code.qualname}: {name} (opcode {op}) not interpreted")
    raise RuntimeError(f"{code.qualname}: fell off the end of the bytecode")


def opcode_histogram(code: Code) -> dict:
    """Opcode names used by a code object (for reporting what the interpreter had to support)."""
    out: dict[str, int] = {}
    bc = code.code
    for i in range(0, len(bc), 2):
        n = OP.get(bc[i], f"<{bc[i]}>")
        if n != "CACHE":
            out[n] = out.get(n, 0) + 1
    return out
