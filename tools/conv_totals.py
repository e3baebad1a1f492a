"""Sum per-layer conv timings over the UNet(3,150) layer list (multiplicities) for the library and for ours."""
import json, sys
COUNTS = {(64, 64, 128): 2, (128, 128, 128): 2, (128, 64, 128): 1, (64, 64, 64): 2, (64, 128, 64): 1, (128, 128, 64): 1,
          (256, 256, 64): 2, (256, 128, 64): 1, (128, 64, 64): 1, (128, 128, 32): 2, (128, 256, 32): 1, (256, 256, 32): 1,
          (512, 512, 32): 2, (512, 256, 32): 1, (256, 128, 32): 1, (256, 256, 16): 5, (256, 512, 16): 1, (512, 512, 16): 3,
          (512, 256, 16): 1}
for f in sys.argv[1:]:
    tot = {"fwd_ms": 0.0, "dgrad_ms": 0.0, "wgrad_ms": 0.0}
    text = open(f).read()
    try:                                   # tools/with_clocks.py document ({"clocks": ..., "records": [...]})
        doc = json.loads(text)
        recs = doc["records"] if isinstance(doc, dict) and "records" in doc else None
    except ValueError:
        recs = None
    if recs is None:                       # plain JSON lines
        recs = [json.loads(line) for line in text.splitlines() if line.startswith("{")]
    for d in recs:
        n = COUNTS[(d["cin"], d["cout"], d["hw"])]
        for k in tot:
            tot[k] += n * d[k]
    print(f, {k: round(v, 2) for k, v in tot.items()}, "total", round(sum(tot.values()), 2))
