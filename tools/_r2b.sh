python tools/pr_exact_sweep.py 26 "GDN_PR_EXACT_COLS=1000000000;GDN_PR_EXACT_COLS=65536;GDN_PR_EXACT_COLS=32768;GDN_PR_EXACT_COLS=16384;GDN_PR_EXACT_COLS=8192;GDN_PR_EXACT_COLS=16384,GDN_PR_BANDS=0;GDN_PR_EXACT=1"
python - <<'PY'
import sys; sys.path.insert(0,'.')
import torch, bench, gardenia_b200 as gb
pre,g=bench.load_graph("g",26)
dg=gb.DeviceGraph(g)
d=torch.empty(g.m,dtype=torch.int32,device="cuda")
srcs=[int(x) for x in g.pick_sources(4)]
dg.bfs(srcs[0],d)
for s in srcs:
    st=dg.bfs(s,d)
    print("BFS",s,"ms",round(st.solve_ms,3),"bu_ms",round(st.kernel_ms,3),[(x["dir"],x["ns"]//1000,x["frontier"],x["edges"]) for x in st.bfs_steps()])
PY
