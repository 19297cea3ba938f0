import sys,re
tot=[];rounds=[]
for l in sys.stdin:
    m=re.match(r"reset e=(\d+) scn=(\d+) robot (\d+) first-half (\d+) second-half (\d+) cycles, (\d+) rounds",l)
    if m:
        tot.append(int(m.group(3))+int(m.group(4))+int(m.group(5))); rounds.append(int(m.group(6)))
tot=tot[-300:]; rounds=rounds[-300:]   # steady-state steps only (the first launch resets every env at once)
tot.sort()
if tot: print("n=%d median %.1f us  p90 %.1f us  max %.1f us  (1.965 GHz); rounds median %d max %d"%(len(tot),tot[len(tot)//2]/1965,tot[int(len(tot)*0.9)]/1965,tot[-1]/1965,sorted(rounds)[len(rounds)//2],max(rounds)))
