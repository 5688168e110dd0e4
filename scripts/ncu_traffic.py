"""Join an `ncu --set full` capture (raw CSV page) of the gather-GEMM launches of one forward with
gpurun_out/unet_launches.json (scripts/prof_unet_step.py) -> profiles/<out>.json: per launch the kernel ncu saw, its duration,
DRAM bytes read + written, L2 bytes (lts__t_sectors * 32), tensor-pipe and DRAM utilisation, next to the algorithmic bytes.

    python scripts/ncu_traffic.py gpurun_out/prof_unet_raw.csv gpurun_out/unet_launches.json profiles/r02_gather_gemm_traffic.json
"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
        "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def main(raw_csv, launches_json, out_json, note=""):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, default=None):
        i = col.get(name)
        if i is None or r[i] in ("", "n/a"):
            return default
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    L = json.load(open(launches_json))
    out = []
    for k, r in enumerate(data):
        if k >= len(L["launches"]):
            break
        d = L["launches"][k]
        rd, wr = val(r, "dram__bytes_read.sum", 0.0), val(r, "dram__bytes_write.sum", 0.0)
        sect = val(r, "lts__t_sectors.sum")
        out.append(dict(name=f"launch {k}: {'sparse' if d['sparse'] else 'dense'} koff={d['koff']} rows_in={d['rows_in']} "
                             f"m_out={d['m_out']} {d['cin']}->{d['cout']}", kernel=r[col["Kernel Name"]][:80],
                        grid=r[col["Grid Size"]], duration_us=val(r, "gpu__time_duration.sum"), dram_read_bytes=rd,
                        dram_write_bytes=wr, traffic_bytes=rd + wr, l2_bytes=None if sect is None else sect * 32.0,
                        tensor_pipe_pct=val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                        dram_pct=val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        issue_active_pct=val(r, "sm__inst_issued.avg.pct_of_peak_sustained_active"),
                        algorithmic_bytes=d["algorithmic_bytes"], pairs=d["pairs"], flops=d["flops"],
                        event_us_outside_ncu=d["event_us"], note=note))
    sp = [o for o in out if o["name"].split(": ")[1].startswith("sparse")]
    # the bench line quotes launches[0]: put the largest sparse launch (a level-2 SubM conv) first
    sp.sort(key=lambda o: -o["algorithmic_bytes"])
    rest = [o for o in out if o not in sp]
    json.dump(dict(source=f"ncu --set full --clock-control none of {L['engine']} inside one eager MSeg3D forward "
                          f"(scripts/prof_unet_step.py); dram__bytes_read.sum + dram__bytes_write.sum, lts__t_sectors.sum * 32 "
                          f"per launch (cold-cache, serialised replays)", launches=sp + rest), open(out_json, "w"), indent=1)
    for o in (sp + rest)[:12]:
        print(o["name"], "|", o["kernel"][:40], f"{o['duration_us']:.1f} us dram {o['traffic_bytes'] / 1e6:.1f} MB "
              f"alg {o['algorithmic_bytes'] / 1e6:.1f} MB l2 {(o['l2_bytes'] or 0) / 1e6:.1f} MB tensor {o['tensor_pipe_pct']}")


if __name__ == "__main__":
    main(*sys.argv[1:])
