"""Fuzzer of the forcing disaggregation alone (CPU): the reference in its disaggregator mode (OUTPUT_FORCE TRUE, oracle/_ref/vic_ref_harness) against
the host build of vic_disagg.cuh (oracle/_ref/disaggport) on random places of the globe -- polar circles and poles (polar night, midnight sun), the
equator, the southern hemisphere; longitudes up to 20 hours away from the model's time zone, in both directions and off the whole hour -- start days
through the year, leap years, 1- / 3- / 6-hourly steps and the disaggregation's options; the hourly forcing must be identical bit for bit.
  python tools/disagg_fuzz.py [seed] [trials]
(this is the fuzzer that found the one summation-order difference of the round: an hour that holds both the last and the first 30 s slots of the solar
day, mtclim_wrapper.c:240-249, visible only under the midnight sun)"""
import sys, os, subprocess, dataclasses, numpy as np, tempfile, shutil, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vic_b200 import synth
from vic_b200.casefile import read_case, write_case
from vic_b200.layout import TABLES
REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref')
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 1)
bad=0; n=0
for trial in range(int(sys.argv[2]) if len(sys.argv)>2 else 40):
    lat0=float(rng.choice([66.03125, 68.03125, 72.03125, 80.03125, -70.03125, 89.03125, 60.03125, 0.03125]))
    lon0=float(rng.choice([-150.03125,-142.53125,-135.03125,-127.53125,-121.96875,-112.53125,-100.03125,-90.03125, -120.03125, -119.96875, 10.03125, 179.03125]))
    startday=int(rng.choice([1,80,150,172,182,200,265,330,355]))
    year=int(rng.choice([2001,2004]))
    dt=int(rng.choice([1,1,3,6,24]))
    snow_step=dt if dt < 24 else int(rng.choice([1,3,6]))  # daily model step: the sub-daily snow step sets the forcing slots (NF = 24 / SNOW_STEP)
    extra=[f"LW_TYPE {rng.choice(['LW_TVA','LW_PRATA','LW_IDSO'])}", f"VP_ITER {rng.choice(['VP_ITER_ALWAYS','VP_ITER_NONE','VP_ITER_CONVERGE','VP_ITER_ANNUAL'])}", f"VP_INTERP {rng.choice(['TRUE','FALSE'])}", f"PLAPSE {rng.choice(['TRUE','FALSE'])}"]
    cfg=dataclasses.replace(synth.CONFIGS["disagg"], ndays=int(rng.integers(5,40)), startday=startday, startyear=year, dt=dt, snow_step=snow_step, extra_global=extra)
    d=tempfile.mkdtemp(prefix='polar_')
    try:
        r=synth.generate(d+'/in', cfg, 2, 3, int(rng.integers(1,1<<30)), lat0=lat0, lon0=lon0)
        h=subprocess.run([f'{REF}/vic_ref_harness','-g',r['global_file'],'-o',d+'/case.bin'],capture_output=True,text=True)
        if h.returncode!=0: print('skip (reference rc',h.returncode,')',lat0,lon0,startday); continue
        c=read_case(d+'/case.bin')
        lat,lng=(c["cellpar"][:, TABLES["cpar"].index(k)] for k in ("CP_lat","CP_lng"))
        nd=int(c["disagg_raw"][5])
        daily=np.stack([np.loadtxt(os.path.join(r["dir"],"forc",f"f_{la:.5f}_{lo:.5f}"))[:nd] for la,lo in zip(lat,lng)])
        write_case(d+'/d.bin',{"options_raw":c["options_raw"],"disagg_raw":c["disagg_raw"],"meta":c["meta"],"cellpar":c["cellpar"],"daily":daily})
        q=subprocess.run([f'{REF}/disaggport',d+'/d.bin',d+'/f.bin'],capture_output=True,text=True)
        if q.returncode!=0: print('PORT ERROR', q.stderr[-200:]); bad+=1; continue
        f=read_case(d+'/f.bin')["forcing"]; n+=1
        same=np.array_equal(f,c["forcing"])
        if not same:
            bad+=1; ne=(f!=c["forcing"]); print("DIFF", lat0, lon0, startday, year, dt, snow_step, extra, ne.sum(), sorted(set(np.nonzero(ne)[2])))
    finally:
        shutil.rmtree(d, ignore_errors=True)
print("compared", n, "different", bad)
sys.exit(1 if bad else 0)
