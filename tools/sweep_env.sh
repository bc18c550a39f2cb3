for ts in 0 5 10 20; do echo "== TAIL_SPLIT=$ts"; DFA_FWD_TAIL_SPLIT=$ts python tools/sweep_fwd.py 30,33 rig:8:f32,rig:4:f32; done
