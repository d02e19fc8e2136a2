"""Runs the UNMODIFIED reference (zhuker/lamejs, /root/reference) under a real JavaScript engine: Qt's QJSEngine
(libQt6Qml 6.6.3), which ships inside the Nsight Compute host directory of this image.  Test infrastructure only --
it exists to pin oracle/ (and through it the CUDA path) against lamejs itself: tests/golden/make_lamejs_golden.py
records lamejs's own output bytes as committed fixtures; tests/test_lamejs_pin.py checks the oracle against them
(and live against the engine when /root/reference is present).

Two ways of loading the reference are supported and must agree:
  bundle  = /root/reference/lame.all.js           (the reference's own concatenation, makeall.sh)
  modules = /root/reference/src/js/*.js           (CommonJS sources through a 12-line `require` shim)
Math.* is the engine's (QV4 -> C libm).  `fdlibm=True` swaps Math.log10/log/exp/pow/sin/cos/atan for a JavaScript
transcription of fdlibm 5.3 (what V8's base/ieee754 ports), see fdlibm.js, to show the bytes do not depend on it."""
import glob
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
RUNNER = os.path.join(BUILD, "qjs_run")
REF = "/root/reference"


def qt_dir():
    c = sorted(glob.glob("/opt/nvidia/nsight-compute/*/host/linux-desktop-glibc_2_11_3-x64"))
    return c[-1] if c else None


def available():
    q = qt_dir()
    return bool(q) and os.path.exists(os.path.join(q, "libQt6Qml.so.6")) and os.path.exists(os.path.join(REF, "lame.all.js"))


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def _env():
    e = dict(os.environ)
    e["QT_NO_GLIB"] = "1"
    e["LC_ALL"] = "C.UTF-8"
    e["QT_LOGGING_RULES"] = "qt.qml.compiler=false"
    e["LD_LIBRARY_PATH"] = BUILD + ":" + qt_dir() + ":" + e.get("LD_LIBRARY_PATH", "")
    return e


def run_js(files, timeout=3600):
    """Evaluates the JS files in order in one engine; returns the last completion value as a string."""
    if not os.path.exists(RUNNER):
        build()
    p = subprocess.run([RUNNER] + list(files), env=_env(), capture_output=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError("qjs_run failed: " + p.stderr.decode(errors="replace")[-2000:])
    return p.stdout.decode()


_MODULE_ORDER = None


def modules_loader_source(hooks=False):
    """A CommonJS loader for /root/reference/src/js: every file becomes a factory function in a table; `require`
    instantiates on first use (cycles see the partially filled exports, as in node)."""
    src_dir = os.path.join(REF, "src", "js")
    parts = ["var __factories = {}; var __cache = {};\n"
             "function __require(name){ name = name.replace(/^\\.\\//,''); if(!/\\.js$/.test(name)) name += '.js';\n"
             "  if(__cache[name]) return __cache[name].exports; var m = {exports:{}}; __cache[name] = m;\n"
             "  if(!__factories[name]) throw new Error('module not found: '+name);\n"
             "  __factories[name](m, m.exports, __require); return m.exports; }\n"]
    for f in sorted(os.listdir(src_dir)):
        if not f.endswith(".js") or f == "Tests.js":
            continue
        body = open(os.path.join(src_dir, f), encoding="utf-8", errors="replace").read()
        if hooks:
            for hf, anchor, ins in _HOOKS:
                if hf == f:
                    assert body.count(anchor) == 1, (f, anchor)
                    body = body.replace(anchor, ins + anchor)
        parts.append("__factories[%s] = function(module, exports, require){\n%s\n};\n" % (json.dumps(f), body))
    parts.append("var lamejs = __require('index.js');\n")
    return "".join(parts)


_HOOKS = [   # (file, anchor text, inserted before the anchor) -- patched into the in-memory module text only
    ("Encoder.js", "gfc.iteration_loop.iteration_loop(gfp, pe_use, ms_ener_ratio, masking);",
     "if (typeof __lj_pre === 'function') __lj_pre(gfp, masking);\n        "),
    ("Encoder.js", "/* copy mp3 bit buffer into array */\n        mp3count = bs.copy_buffer(gfc, mp3buf, mp3bufPos, mp3buf_size, 1);",
     "if (typeof __lj_post === 'function') __lj_post(gfp);\n        "),
]

_TAPS_JS = r"""
var __taps = [];
var __f32 = new Float32Array(1), __u32 = new Uint32Array(__f32.buffer);
function __bits(a, n) { var o = []; for (var i = 0; i < n; i++) { __f32[0] = a[i]; o.push(__u32[0]); } return o; }
function __lj_pre(gfp, masking) {
  var gfc = gfp.internal_flags, fr = {xr: [], en_l: [], thm_l: [], en_s: [], thm_s: [], bt: [], ath: gfc.ATH.adjust};
  for (var gr = 0; gr < gfc.mode_gr; gr++) for (var ch = 0; ch < gfc.channels_out; ch++) {
    var gi = gfc.l3_side.tt[gr][ch], m = masking[gr][ch];
    fr.xr.push(__bits(gi.xr, 576)); fr.bt.push(gi.block_type);
    fr.en_l.push(__bits(m.en.l, 22)); fr.thm_l.push(__bits(m.thm.l, 22));
    var es = [], ts = [];
    for (var sb = 0; sb < 13; sb++) { es = es.concat(__bits(m.en.s[sb], 3)); ts = ts.concat(__bits(m.thm.s[sb], 3)); }
    fr.en_s.push(es); fr.thm_s.push(ts);
  }
  __taps.push(fr);
}
function __lj_post(gfp) {
  var gfc = gfp.internal_flags, fr = __taps[__taps.length - 1];
  fr.l3 = []; fr.gg = []; fr.p23 = []; fr.p2 = []; fr.bv = []; fr.c1 = []; fr.sfc = []; fr.sf = [];
  for (var gr = 0; gr < gfc.mode_gr; gr++) for (var ch = 0; ch < gfc.channels_out; ch++) {
    var gi = gfc.l3_side.tt[gr][ch];
    fr.l3.push(Array.prototype.slice.call(gi.l3_enc, 0, 576)); fr.gg.push(gi.global_gain); fr.p23.push(gi.part2_3_length);
    fr.p2.push(gi.part2_length); fr.bv.push(gi.big_values); fr.c1.push(gi.count1); fr.sfc.push(gi.scalefac_compress);
    fr.sf.push(Array.prototype.slice.call(gi.scalefac, 0, 39));
  }
}
"""


def _hex16(a):
    return np.ascontiguousarray(a, dtype="<i2").tobytes().hex()


_DRIVER = r"""
function __unhex(h){ var n=h.length/4; var a=new Int16Array(n); for(var i=0;i<n;i++){
  var v=parseInt(h.substr(4*i+2,2)+h.substr(4*i,2),16); a[i]= v>=32768? v-65536: v;} return a;}
function __tohex(b){ var s=[]; for(var i=0;i<b.length;i++){ var v=b[i]&255; s.push((v<16?"0":"")+v.toString(16)); } return s.join(""); }
(function(){
  var L=__unhex(__HEXL), R=__unhex(__HEXR);
  var t0=Date.now();
  var enc = new lamejs.Mp3Encoder(__CH, __SR, __KBPS);
  var t1=Date.now();
  var parts=[], n=L.length, step=__CHUNK>0?__CHUNK:Math.max(n,1);
  for (var i=0;i<n;i+=step) parts.push(__CH==2 ? enc.encodeBuffer(L.subarray(i,i+step), R.subarray(i,i+step))
                                              : enc.encodeBuffer(L.subarray(i,i+step)));
  parts.push(enc.flush());
  var t2=Date.now();
  var sizes=[], hex=[];
  for (var p=0;p<parts.length;p++){ sizes.push(parts[p].length); hex.push(__tohex(parts[p])); }
  return JSON.stringify({sizes:sizes, init_ms:t1-t0, encode_ms:t2-t1, hex:hex.join("")});
})();
"""


def encode(channels, samplerate, kbps, left, right=None, chunk=None, loader="bundle", fdlibm=False, extra_js=None, driver=None):
    """new lamejs.Mp3Encoder(channels, samplerate, kbps); encodeBuffer(whole stream or `chunk`-sample calls); flush().
    Returns (bytes, per-call sizes, info dict)."""
    if right is None:
        right = left
    with tempfile.TemporaryDirectory() as td:
        files = []
        if fdlibm:
            files.append(os.path.join(HERE, "fdlibm.js"))
        if loader == "bundle":
            files.append(os.path.join(REF, "lame.all.js"))
        else:
            p = os.path.join(td, "modules.js")
            open(p, "w").write(modules_loader_source(hooks=(loader == "taps")))
            files.append(p)
        if extra_js:
            p = os.path.join(td, "extra.js")
            open(p, "w").write(extra_js)
            files.append(p)
        d = os.path.join(td, "drive.js")
        with open(d, "w") as f:
            f.write('var __HEXL="%s"; var __HEXR="%s"; var __CH=%d, __SR=%d, __KBPS=%d, __CHUNK=%d;\n'
                    % (_hex16(left), _hex16(right), channels, samplerate, kbps, chunk or 0))
            f.write(driver or _DRIVER)
        files.append(d)
        o = json.loads(run_js(files))
    data = bytes.fromhex(o.pop("hex"))
    return data, o["sizes"], o


def encode_with_taps(channels, samplerate, kbps, left, right=None, chunk=None):
    """Like encode(), with lamejs's own intermediates of every frame (module sources + two one-line hooks patched into the
    in-memory text of Encoder.js): returns (bytes, taps) where taps maps name -> numpy array shaped like oracle_lib's trace:
    xr/en_l/thm_l/en_s/thm_s as float32 bit patterns, blocktype, ath_adjust, l3_enc, global_gain, part2_3_length, ..."""
    drv = _DRIVER.replace("return JSON.stringify({sizes:sizes,", "return JSON.stringify({taps:__taps, sizes:sizes,")
    data, sizes, o = encode(channels, samplerate, kbps, left, right, chunk=chunk, loader="taps", extra_js=_TAPS_JS, driver=drv)
    fr = o["taps"]
    F = len(fr)
    ngc = len(fr[0]["bt"]) if F else 0
    G = ngc // channels if F else 0

    def arr(key, dt, inner):
        a = np.array([f[key] for f in fr], dtype=dt)
        return a.reshape((F, G, channels) + inner)

    taps = {
        "xr": arr("xr", np.uint32, (576,)).view(np.float32), "en_l": arr("en_l", np.uint32, (22,)).view(np.float32),
        "thm_l": arr("thm_l", np.uint32, (22,)).view(np.float32), "en_s": arr("en_s", np.uint32, (13, 3)).view(np.float32),
        "thm_s": arr("thm_s", np.uint32, (13, 3)).view(np.float32), "blocktype": arr("bt", np.int32, ()),
        "ath_adjust": np.array([f["ath"] for f in fr], dtype=np.float64), "l3_enc": arr("l3", np.int32, (576,)),
        "global_gain": arr("gg", np.int32, ()), "part2_3_length": arr("p23", np.int32, ()), "part2_length": arr("p2", np.int32, ()),
        "big_values": arr("bv", np.int32, ()), "count1": arr("c1", np.int32, ()), "scalefac_compress": arr("sfc", np.int32, ()),
        "scalefac": arr("sf", np.int32, (39,)),
    }
    return data, taps
