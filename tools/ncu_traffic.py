"""Per-kernel DRAM traffic and key counters of an .ncu-rep, as JSON + text.
  python tools/ncu_traffic.py file.ncu-rep <batch> out.json out.txt
The LAST launch of every kernel name in the report is used (the captures run two iterations; the second one is warm)."""
import csv, io, json, subprocess, sys

LABEL = [("unstuff_count", "jpeg_unstuff_count"), ("unstuff_scan", "jpeg_unstuff_scan"), ("unstuff_scatter", "jpeg_unstuff_scatter"),
         ("huff_sync_intra", "jpeg_huff_sync_intra"), ("huff_sync_walk", "jpeg_huff_sync_walk"), ("huff_scan", "jpeg_huff_scan"),
         ("huff_write", "jpeg_huff_write"), ("dc_scan", "jpeg_dc_scan"), ("truncation_fixup", "jpeg_truncation_fixup"), ("idct_kernel", "jpeg_idct"),
         ("color_fast", "jpeg_upsample_color"), ("color_kernel", "jpeg_upsample_color_generic"), ("jpeg_post", "jpeg_post"),
         ("resample_stream", "resample_stream"), ("resample_planar", "resample_planar"), ("resample_fused", "resample_fused"),
         ("cmn_hwc2chw", "cmn_hwc2chw"), ("cmn_generic", "cmn_generic"), ("warp_affine_tma", "warp_affine_tma"), ("warp_affine", "warp_affine"),
         ("linear_transform", "linear_transform"), ("spectrogram1024_kernel<(bool)1>", "spectrogram_mel_fused"), ("spectrogram1024", "spectrogram_stft"),
         ("spectrogram_kernel", "spectrogram_stft_radix2"), ("mel_mma", "mel_filter_bank_mma"), ("mel_kernel", "mel_filter_bank")]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_op_hmma.sum', 'lts__t_sector_hit_rate.pct']


def to_bytes(v, unit):
    v = float(v.replace(",", "") or 0)
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def to_us(v, unit):
    v = float(v.replace(",", "") or 0)
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6, "s": 1e6}.get(unit.lower(), 1)


def main():
    rep, batch, out_json, out_txt = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
    allrows, order = {}, []
    for r in rows[2:]:
        name = r[ki]
        label = next((lab for key, lab in LABEL if key in name), name[:40])
        if label not in allrows:
            order.append(label)
        allrows.setdefault(label, []).append(r)
    # the synchronisation walk is launched three times per batch (walk1 / walk2 / walk3): keep the last three launches apart
    last, order2 = {}, []
    for label in order:
        rs = allrows[label]
        if label == "jpeg_huff_sync_walk" and len(rs) >= 3:
            for k in range(3):
                last[f"{label}{k + 1}"] = rs[len(rs) - 3 + k]
                order2.append(f"{label}{k + 1}")
        else:
            last[label] = rs[-1]
            order2.append(label)
    order = order2
    res, txt = {}, []
    for label in order:
        r = last[label]
        g = lambda m: (r[hdr.index(m)], units[hdr.index(m)]) if m in hdr else ("0", "")
        rd, wr = to_bytes(*g('dram__bytes_read.sum')), to_bytes(*g('dram__bytes_write.sum'))
        res[label] = {"dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
                      "ncu_duration_us": to_us(*g('gpu__time_duration.sum')), "kernel": r[ki][:120]}
        txt.append(f"---- {label}: {r[ki][:110]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                txt.append(f"  {w[:62]:62s} {r[i][:22]} {units[i]}")
        st = sorted(((float((r[hdr.index(h)] or '0').replace(',', '')), h[34:-24]) for h in stalls), reverse=True)[:6]
        txt.append('  stalls/issue: ' + ' '.join(f'{n}={v:.2f}' for v, n in st))
    json.dump({"batch": batch, "source": f"ncu --set full --clock-control none, last launch of every kernel in {rep.split('/')[-1]}",
               "kernels": res}, open(out_json, "w"), indent=1)
    open(out_txt, "w").write("\n".join(txt) + "\n")
    print("\n".join(txt))


if __name__ == "__main__":
    main()
