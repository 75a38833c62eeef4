"""Back-of-the-envelope pipe accounting for the hd128 attention kernel and its prepared variants (no GPU needed).

Inputs are the measured figures of profiles/r01_ncu_full_summary.txt (3.49 ms, 1.786 GHz, 1.977e9 warp instructions, 8192 CTAs x
128 tiles of 128 x 64) and the per-variant instruction deltas read from the SASS.  Per SM and per PAIR of
tiles (one per resident CTA) each of the four schedulers has
    MUFU   : 2 tiles x 32 rows x 64 exponentials / 4 lanes per clock          = 1024 cycles
    issue  : warp instructions of its 2 softmax + 2 correction warps (+ TMA / MMA share)
    tensor : 2 x (QK 128 + PV 128) cycles of the SM-wide pipe                 =  512 cycles
and the kernel cannot be faster than the largest of them; today it sits at 1759 cycles because none of them is the limiter —
the serial chain of a softmax warp is.  The script prints, per variant, the pipe bounds and the throughput they would allow, i.e.
what is to gain at most once the chain is hidden (more chains per scheduler: attn_alt.cu) or shortened (PREMAX / DEFER / LATE_ALPHA).
"""
SMS, GHZ = 148, 1.786
TILE_FLOP = 4.0 * 128 * 64 * 128
MEASURED_PAIR_CYCLES = 3.49e-3 * GHZ * 1e9 / (8192 * 128 / SMS / 2)
ISSUE_NOW = 1.977e9 / (SMS * 4) / (8192 * 128 / SMS / 2)          # warp instructions per scheduler and tile pair


def pflops(pair_cycles):
    return SMS * 2 * TILE_FLOP * GHZ * 1e9 / pair_cycles / 1e15


VARIANTS = [
    # name, MUFU cycles per pair and scheduler, issue slots per pair and scheduler, note
    ("product (measured)", 1024, ISSUE_NOW, "chain-bound: 1759 cycles per pair"),
    ("lazy3", 1024, ISSUE_NOW - 2 * 0.8 * 90, "correction warps' rescale (~90 instructions) in ~3 % instead of ~80 % of the warp-tiles"),
    ("poly1", 768, ISSUE_NOW + 2 * 64, "16 of 64 exponentials per row on the FMA pipe: -256 MUFU cycles, +64 slots per tile"),
    ("lazy3 + poly1", 768, ISSUE_NOW - 2 * 0.8 * 90 + 2 * 64, ""),
    ("poly2", 512, ISSUE_NOW + 2 * 128, "32 of 64"),
    ("alt (two softmax warpgroups per CTA)", 1024, ISSUE_NOW + 2 * 40 - 2 * 0.8 * 90 * 0, "same work + m hand-off + two-pass TMEM loads (~40 per tile); in-line rescale"),
    ("alt + lazy3", 1024, ISSUE_NOW + 2 * 40 - 2 * 0.8 * 90, ""),
    ("alt + lazy3 + poly1", 768, ISSUE_NOW + 2 * 40 - 2 * 0.8 * 90 + 2 * 64, ""),
]

if __name__ == "__main__":
    print(f"measured: {MEASURED_PAIR_CYCLES:.0f} cycles per tile pair and SM = {pflops(MEASURED_PAIR_CYCLES):.2f} PFLOP/s; "
          f"issue {ISSUE_NOW:.0f} slots, MUFU 1024, tensor 512 per scheduler")
    print(f"{'variant':40s} {'MUFU':>6s} {'issue':>6s} {'tensor':>6s} {'bound':>6s} {'PFLOP/s at the bound':>22s}")
    for name, mufu, issue, note in VARIANTS:
        bound = max(mufu, issue, 512)
        print(f"{name:40s} {mufu:6.0f} {issue:6.0f} {512:6d} {bound:6.0f} {pflops(bound):22.2f}   {note}")
