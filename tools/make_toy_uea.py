"""Write a small UEA-format archive (<out>/<name>/<name>_{TRAIN,TEST}.ts) from the synthetic class-conditional generator,
with variable series lengths and a few missing samples, to exercise the real-data path of run.py:
    python tools/make_toy_uea.py /tmp/uea Toy && python speech-imagery-eeg_b200/run.py --data UEA --data_root /tmp/uea --dataset Toy ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "speech-imagery-eeg_b200"))
import torch  # noqa: E402
from data_provider.data_factory import SyntheticSeries  # noqa: E402


def write(path, ds, gen):
    with open(path, "w") as f:
        f.write("@problemName Toy\n@timeStamps false\n@missing true\n@univariate false\n@dimensions %d\n@equalLength false\n"
                "@classLabel true %s\n@data\n" % (ds.enc_in, " ".join("c%d" % c for c in range(ds.num_class))))
        for i in range(len(ds)):
            x, y = ds[i]
            n = int(torch.randint(int(0.8 * x.shape[0]), x.shape[0] + 1, (1,), generator=gen))
            dims = []
            for c in range(x.shape[1]):
                vals = ["%.6f" % v for v in x[:n, c].tolist()]
                if i % 7 == 0 and c == 0:
                    vals[n // 2] = "?"
                dims.append(",".join(vals))
            f.write(":".join(dims) + ":c%d\n" % int(y))


if __name__ == "__main__":
    out, name = sys.argv[1], sys.argv[2]
    os.makedirs(os.path.join(out, name), exist_ok=True)
    g = torch.Generator().manual_seed(0)
    write(os.path.join(out, name, name + "_TRAIN.ts"), SyntheticSeries(6, 60, 4, 160, 11), g)
    write(os.path.join(out, name, name + "_TEST.ts"), SyntheticSeries(6, 60, 4, 64, 33), g)
    print("wrote", os.path.join(out, name))
