#!/usr/bin/env python
"""Regenerates the ctypes mirror of `ffvc_gemm_params` in INTEGRATION.md from include/ffvc.h (between the BEGIN / END markers),
so the documented stub cannot drift from the header; tests/test_abi.py executes the documented block and compares its
ctypes.sizeof with ffvc_sizeof("ffvc_gemm_params") and its field order with the package's own mirror."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CT = {"const void*": "ctypes.c_void_p", "void*": "ctypes.c_void_p", "const float*": "ctypes.c_void_p", "int": "ctypes.c_int",
      "int64_t": "ctypes.c_int64", "float": "ctypes.c_float"}


def struct_fields():
    src = open(os.path.join(ROOT, "include", "ffvc.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} ffvc_gemm_params;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", " ", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"^((?:const )?\w+\*?)\s+(.*)$", decl)
        ctype = m.group(1)
        for name in m.group(2).split(","):
            name = name.strip()
            t = ctype
            if name.startswith("*"):
                t, name = ctype + "*", name[1:]
            fields.append((name, CT[t]))
    return fields


def stub():
    f = struct_fields()
    lines = ["class ffvc_gemm_params(ctypes.Structure):            # generated from include/ffvc.h by tools/gen_ctypes_stub.py",
             "    _fields_ = ["]
    row = "        "
    for name, t in f:
        item = '("%s", %s), ' % (name, t)
        if len(row) + len(item) > 118:
            lines.append(row.rstrip())
            row = "        "
        row += item
    lines.append(row.rstrip().rstrip(",") + "]")
    return "\n".join(lines)


def main():
    p = os.path.join(ROOT, "INTEGRATION.md")
    s = open(p).read()
    new = "<!-- BEGIN GENERATED STRUCT -->\n```python\nimport ctypes\n" + stub() + "\n```\n<!-- END GENERATED STRUCT -->"
    s = re.sub(r"<!-- BEGIN GENERATED STRUCT -->.*?<!-- END GENERATED STRUCT -->", lambda m: new, s, flags=re.S)
    open(p, "w").write(s)
    print(new)


if __name__ == "__main__":
    main()
