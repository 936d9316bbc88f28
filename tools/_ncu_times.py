import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
for r in rows:
    name = r[4].split("(")[0][-60:]
    print(f"{name:60s} {r[-1]:>14s} {r[-2]}")
