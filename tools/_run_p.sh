timeout 120 python tools/pcheck.py
SFH_FLAT=1 timeout 120 python tools/pcheck.py
SFH_FLAT=1 timeout 300 python tools/kbench.py --only train_ --hd
SFH_FLAT=1 timeout 120 python tools/kbench.py --only train_640x --oob
