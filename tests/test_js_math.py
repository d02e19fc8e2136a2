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
