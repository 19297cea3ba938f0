"""stdin: the kernel list of tools/pipeline_timeline.py; prints the time between successive crowd-step kernel starts (= one rollout step)."""
import sys

ts=[float(l.split('us')[0]) for l in sys.stdin if 'crowd_step_kernel' in l]
d=[b-a for a,b in zip(ts,ts[1:])]
print("step periods us:", ["%.1f"%x for x in d])
