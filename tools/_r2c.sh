export GDN_PR_KTIME=1
python tools/pr_exact_sweep.py 26 "GDN_PR_EXACT_COLS=65536;GDN_PR_EXACT_BUDGET=34000000;GDN_PR_EXACT_BUDGET=400000000" 2>&1 | grep -v "^\[bench\]" | cut -c1-330 | awk '/iteration 3/{last=$0; next} {if(last!="")print last; last=""; print}'
