import sys; sys.path.insert(0,'.')
import torch, bench, gardenia_b200 as gb
scale=int(sys.argv[1]) if len(sys.argv)>1 else 26
pre,g=bench.load_graph("g",scale)
dg=gb.DeviceGraph(g)
d=torch.empty(g.m,dtype=torch.int32,device="cuda")
srcs=[int(x) for x in g.pick_sources(4)]
dg.bfs(srcs[0],d)
for s in srcs:
    st=dg.bfs(s,d)
    print("BFS",s,"ms",round(st.solve_ms,3),"bu_ms",round(st.kernel_ms,3),[(x["dir"],x["ns"]//1000,x["frontier"],x["edges"]) for x in st.bfs_steps()])
