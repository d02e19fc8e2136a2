"""lamejs in the modes `Mp3Encoder` hard-codes away: the same module wiring as src/js/index.js:66-115,
with gfp.mode / gfp.disable_reservoir chosen by the caller (index.js:104,108 fix them to STEREO / true).  The reference
sources are unmodified; only the driver differs.  Test infrastructure (pins oracle/ for these modes)."""
import json
import os
import tempfile

import ref_lamejs as R
import tag_probe as T

_DRIVER = (T._TAG_DRIVER
           .replace("gfp.mode=MPEGMode.STEREO;", "gfp.mode=MPEGMode[__MODE_NAME];")
           .replace("gfp.bWriteVbrTag=true; gfp.disable_reservoir=true;", "gfp.bWriteVbrTag=false; gfp.disable_reservoir=__NORES;"))


def encode(channels, samplerate, kbps, left, right=None, chunk=None, mode="STEREO", disable_reservoir=True, fdlibm=False):
    """Returns (bytes, per-call sizes, info dict).  fdlibm: load fdlibm.js first (Math.log / log10 / exp / pow as under V8)."""
    if right is None:
        right = left
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "modules.js")
        open(p, "w").write(R.modules_loader_source(hooks=False))
        d = os.path.join(td, "drive.js")
        with open(d, "w") as f:
            f.write('var __HEXL="%s"; var __HEXR="%s"; var __CH=%d, __SR=%d, __KBPS=%d, __CHUNK=%d;\n'
                    % (R._hex16(left), R._hex16(right), channels, samplerate, kbps, chunk or 0))
            f.write("var __MODE_NAME='%s'; var __NORES=%s;\n" % (mode, "true" if disable_reservoir else "false"))
            f.write(_DRIVER)
        o = json.loads(R.run_js(([os.path.join(os.path.dirname(os.path.abspath(__file__)), "fdlibm.js")] if fdlibm else []) + [p, d]))
    data = bytes.fromhex(o.pop("hex"))
    return data, o["sizes"], o
