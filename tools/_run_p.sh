echo default; timeout 300 python tools/kbench.py --hd --only predict
echo NO_REMAP; SFH_NO_REMAP=1 timeout 300 python tools/kbench.py --hd --only predict
echo NO_REMAP store;  SFH_NO_REMAP=1 timeout 300 python tools/kbench.py --hd --only store
