for r in 4 8 12; do echo "B 1 reserve $r"; MKID_K4_RESERVE=$r python scripts/strong_probe.py 200 | grep "pipelined 1"; done
for b in 2 4; do for r in 4 8 12; do echo "B $b reserve $r"; PROBE_B=$b MKID_K4_RESERVE=$r python scripts/strong_probe.py 100 | grep "pipelined 1"; done; done
