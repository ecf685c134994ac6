import re, sys
txt = open(sys.argv[1]).read()
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for sec in txt.split('== ')[1:]:
    name = sec.split('\n')[0]
    ev = [(int(a), int(b)) for a, b in re.findall(r'(\d+)@(-?\d+)', sec)]
    start = {"MMA": 90, "EPI": 380 if "380@" in txt else 400, "PROD": 300}[name]
    if len(sys.argv) > 3 and sys.argv[3] == "raw":
        print("==", name)
        print(" ".join(f"{t}:{c - ev[0][1]}" for t, c in ev[:int(sys.argv[4]) if len(sys.argv) > 4 else 120]))
        continue
    idx = [i for i, (t, _) in enumerate(ev) if t == start]
    if name == "PROD":
        idx = idx[::2]
    if len(idx) <= tile + 1:
        continue
    a, b = idx[tile], idx[tile + 1]
    print("==", name, "abs start", ev[a][1])
    print(" ".join(f"{t}:{c - ev[a][1]}" for t, c in ev[a:b + 1]))
