#!/usr/bin/env python
"""Fuzzes the command-line handling of the ACFcalculator front-end against the shipped reference binary
(oracle/_ref/ACFcalculator_ref; build container only): random option soups, compared on exit status, first stderr line
and stdout.  Uses the oracle-backed build of the front-end (tests/_build/ACFcalculator_oracle, made by the CLI tests), so
no GPU is involved.  Command lines that use this repo's extensions (-r/--factor, --fft, --batch, --gpus) are skipped.

    python tools/fuzz_options.py [seed] [runs]        -> prints every distinct disagreement, then a count"""
import os
import subprocess, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = os.path.join(ROOT, 'oracle', '_ref', 'ACFcalculator_ref'); ours = os.environ.get('FUZZ_OURS', os.path.join(ROOT, 'tests', '_build', 'ACFcalculator_oracle'))
random.seed(int(sys.argv[1]) if len(sys.argv)>1 else 1)
longs=["input","type","cutoff","acf","output","timestep","maxcorr","field","p_factor","spring","mass","temperature","help","bogus"]
shorts=list("itcaosmfpkweh")+["x","z","1"]
vals=["1","abc","","-5","1e3","0.5","general","namd","x y","--","-1.5e-3","1000000000000","0x10","1.","+2","-m","-h","--help","--maxcorr","-","2147483648","-2147483649"," 5","5 ","inf","nan","1e400"]
def rand_args():
    a=[]
    for _ in range(random.randint(0,6)):
        kind=random.random()
        if kind<0.35:
            o="--"+random.choice(longs)
            if random.random()<0.3: o=o[:random.randint(3,len(o))]
            if random.random()<0.5: a.append(o+"="+random.choice(vals))
            else:
                a.append(o)
                if random.random()<0.8: a.append(random.choice(vals))
        elif kind<0.8:
            o="-"+random.choice(shorts)
            if random.random()<0.3: a.append(o+random.choice(vals))
            else:
                a.append(o)
                if random.random()<0.8: a.append(random.choice(vals))
        else:
            a.append(random.choice(vals))
    return a
def bad_for_us(args):
    # our deliberate extensions: -r, --factor, --fft, --batch, --gpus
    for t in args:
        if t.startswith("-r") or t.startswith("--fa") or t.startswith("--ff") or t.startswith("--b") or t.startswith("--g"): return True
    return False
bad=0; seen=set(); n=0
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 2000):
    args=rand_args()
    if bad_for_us(args): continue
    n+=1
    ra=subprocess.run([ref]+args,capture_output=True,text=True,cwd='/tmp')
    oa=subprocess.run([ours]+args,capture_output=True,text=True,cwd='/tmp')
    key=(ra.returncode, ra.stderr.splitlines()[:1], ra.stdout); ko=(oa.returncode, oa.stderr.splitlines()[:1], oa.stdout)
    if key!=ko:
        sig=(tuple(ra.stderr.splitlines()[:1])[:1],tuple(oa.stderr.splitlines()[:1])[:1])
        bad+=1
        if sig in seen: continue
        seen.add(sig)
        print(args); print("  ref :",ra.returncode, ra.stderr.splitlines()[:1], repr(ra.stdout[:80])); print("  ours:",oa.returncode, oa.stderr.splitlines()[:1], repr(oa.stdout[:80]))
print("runs",n,"mismatches:",bad,"distinct",len(seen))
