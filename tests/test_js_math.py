"""The fdlibm restatement of Math.log/log10/exp/pow (oracle/js_math.h == lamejs_b200/csrc/mp3_math.cuh) against
glibc: both are <1 ulp accurate, so they must agree within 1 ulp (2 for log10); a wrong constant would show up as
a gross error.  Also checks the two copies (oracle / product) are the same function."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "%(root)s/oracle/js_math.h"
#include "%(root)s/lamejs_b200/csrc/mp3_math.cuh"
#include <stdio.h>
#include <stdlib.h>
static uint64_t st=88172645463325252ull;
static double rnd(){ st^=st<<13; st^=st>>7; st^=st<<17; return (st>>11)*(1.0/9007199254740992.0);}
static long ulp(double a,double b){ int64_t x,y; memcpy(&x,&a,8); memcpy(&y,&b,8); return labs(x-y);}
int main(){ long mx[4]={0,0,0,0}, bad=0;
 for(long i=0;i<2000000;i++){ double x=exp((rnd()-0.5)*80), y=(rnd()-0.5)*40, b=rnd()*100, e=(rnd()-0.5)*20; long u;
  u=ulp(js_log(x),log(x)); if(u>mx[0])mx[0]=u; u=ulp(js_log10(x),log10(x)); if(u>mx[1])mx[1]=u;
  u=ulp(js_exp(y),exp(y)); if(u>mx[2])mx[2]=u; u=ulp(js_pow(b,e),pow(b,e)); if(u>mx[3])mx[3]=u;
  if(js_log(x)!=m3_log(x)||js_log10(x)!=m3_log10(x)||js_exp(y)!=m3_exp(y)||js_pow(b,e)!=m3_pow(b,e)) bad++; }
 printf("%%ld %%ld %%ld %%ld %%ld\n",mx[0],mx[1],mx[2],mx[3],bad);
 printf("%%d %%d %%d\n", js_pow(9,.5)==3.0, js_pow(2,10)==1024.0, js_log10(1000)==3.0);
 return 0; }
'''


def test_fdlibm_ports_agree_with_glibc():
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        open(src, "w").write(SRC % {"root": ROOT})
        exe = os.path.join(d, "t")
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"])
        out = subprocess.check_output([exe]).decode().split("\n")
        mx = [int(v) for v in out[0].split()]
        assert mx[0] <= 1 and mx[1] <= 2 and mx[2] <= 1 and mx[3] <= 1, mx
        assert mx[4] == 0, "oracle and product math headers diverged"
        assert out[1].split() == ["1", "1", "1"]


JS_DRIVER = r"""
(function () {
  var f64 = new Float64Array(1), u32 = new Uint32Array(f64.buffer);
  function bits(x) { f64[0] = x; return ('00000000' + u32[1].toString(16)).slice(-8) + ('00000000' + u32[0].toString(16)).slice(-8); }
  function fromBits(h) { u32[1] = parseInt(h.substr(0, 8), 16); u32[0] = parseInt(h.substr(8, 8), 16); return f64[0]; }
  var out = [], F = Math.__fdlibm, N = Math.__fdlibm.native, maxulp = [0, 0, 0, 0];
  function ulp(a, b) { f64[0] = a; var ah = u32[1], al = u32[0]; f64[0] = b; var d = (ah - u32[1]) * 4294967296 + (al - u32[0]); return Math.abs(d); }
  for (var i = 0; i < __ARGS.length; i++) {
    var a = fromBits(__ARGS[i][0]), b = fromBits(__ARGS[i][1]);
    var r = [F.log(a), F.log10(a), F.exp(b), F.pow(a, b)];
    out.push(bits(r[0]) + bits(r[1]) + bits(r[2]) + bits(r[3]));
    var n = [N.log(a), N.log10(a), N.exp(b), N.pow(a, b)];
    for (var k = 0; k < 4; k++) if (r[k] === r[k] && n[k] === n[k] && isFinite(r[k]) && isFinite(n[k])) maxulp[k] = Math.max(maxulp[k], ulp(r[k], n[k]));
  }
  return JSON.stringify({out: out, maxulp: maxulp, installed: Math.log === F.log && Math.pow === F.pow});
})();
"""

C_SRC = r'''
#include "%(root)s/oracle/js_math.h"
#include <stdio.h>
#include <inttypes.h>
static double fb(uint64_t u){ double x; memcpy(&x,&u,8); return x; }
static uint64_t tb(double x){ uint64_t u; memcpy(&u,&x,8); return u; }
int main(){ uint64_t a,b; while (scanf("%%" SCNx64 " %%" SCNx64, &a, &b) == 2) { double x=fb(a), y=fb(b);
  printf("%%016" PRIx64 "%%016" PRIx64 "%%016" PRIx64 "%%016" PRIx64 "\n", tb(js_log(x)), tb(js_log10(x)), tb(js_exp(y)), tb(js_pow(x,y))); } return 0; }
'''


def test_fdlibm_js_equals_the_c_restatement_bit_for_bit():
    """tools/jsrun/fdlibm.js (what makes the build image's engine compute log / log10 / exp / pow the way V8 does) against
    oracle/js_math.h on random and special arguments: identical bit patterns, NaN for NaN."""
    import json
    import struct
    import sys
    import numpy as np
    import pytest
    sys.path.insert(0, os.path.join(ROOT, "tools", "jsrun"))
    import ref_lamejs
    if not ref_lamejs.qt_dir() or not os.path.exists(os.path.join(ref_lamejs.qt_dir(), "libQt6Qml.so.6")):
        pytest.skip("no JavaScript engine in this environment")
    rng = np.random.default_rng(4)
    xs = list(np.exp((rng.random(6000) - 0.5) * 80)) + list(rng.random(2000) * 100) + [0.0, -0.0, 1.0, -1.0, 2.0, 10.0, 1000.0, 1e-310, 5e-324,
                                                                                     float("inf"), float("-inf"), float("nan"), 0.5, 9.0, -8.0, 1e300]
    ys = list((rng.random(6000) - 0.5) * 40) + list((rng.random(2000) - 0.5) * 20) + [0.0, 0.5, -0.5, 2.0, 3.0, -3.0, 1e-320, 800.0, -800.0,
                                                                                     float("inf"), float("-inf"), float("nan"), 1.0 / 3, 0.75, 1.0, 1e300]
    args = [(struct.pack(">d", float(a)).hex(), struct.pack(">d", float(b)).hex()) for a, b in zip(xs, ys)]
    with tempfile.TemporaryDirectory() as d:
        js = os.path.join(d, "drv.js")
        open(js, "w").write("var __ARGS = %s;\n" % json.dumps(args) + JS_DRIVER)
        o = json.loads(ref_lamejs.run_js([os.path.join(ROOT, "tools", "jsrun", "fdlibm.js"), js]))
        src = os.path.join(d, "t.cpp")
        open(src, "w").write(C_SRC % {"root": ROOT})
        exe = os.path.join(d, "t")
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"])
        want = subprocess.run([exe], input="\n".join("%s %s" % a for a in args).encode(), capture_output=True, check=True).stdout.decode().split()
    assert o["installed"] and len(want) == len(args)
    nan = lambda h: (int(h, 16) & 0x7FF0000000000000) == 0x7FF0000000000000 and (int(h, 16) & 0xFFFFFFFFFFFFF) != 0   # noqa: E731
    bad = []
    for i, (g, w) in enumerate(zip(o["out"], want)):
        for k in range(4):
            a, b = g[16 * k:16 * k + 16], w[16 * k:16 * k + 16]
            if a != b and not (nan(a) and nan(b)):
                bad.append((i, k, args[i], a, b))
    assert not bad, bad[:5]
    # and against the engine's own libm: both are < 1 ulp functions (log10: 2)
    assert o["maxulp"][0] <= 1 and o["maxulp"][1] <= 2 and o["maxulp"][2] <= 1 and o["maxulp"][3] <= 1, o["maxulp"]
