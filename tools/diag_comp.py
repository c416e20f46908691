"""Diagnostics: per-component statistics of the max-flow kernel (rounds, pushes, BFS levels, SM
cycles) for configs 4 and 1, written to gpurun_out/comp_<config>.txt via GDS_DUMP_COMP.
Run from the repo root on a GPU box: python tools/diag_comp.py"""
import os,sys,time
sys.path.insert(0,'.')
import numpy as np, torch
from __graft_entry__ import load_package
pkg=load_package()
import bench
for wname in ('c4','c1'):
    wl=bench.WORKLOADS[wname]
    st,en,_,_=bench.generate(wl,[0],pinned=False)
    s=st.numpy().view(np.uint32); e=en.numpy().view(np.uint32)
    solver=pkg.Solver(0)
    for i in range(2): r=solver.solve(s,e,wl['L'],wl['M'])
    os.environ['GDS_DUMP_COMP']='gpurun_out/comp_%s.txt'%wname
    r=solver.solve(s,e,wl['L'],wl['M'])
    del os.environ['GDS_DUMP_COMP']
    print(wname,'ms_maxflow',r.ms_maxflow,'rounds_max',r.rounds_max,'comps',r.n_components)
    solver.close()
