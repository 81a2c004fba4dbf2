"""Referee for RVQ code parity (index work whose input is floating point)."""
import torch

from oracle import mimi as omimi

TIE_TOL = 1e-4


def assert_codes_match(got, ref, clips, mimi_weights):
    """Codes identical to the reference codes, or — where they differ — a float-level near-tie of the nearest-neighbour search:
    every disagreeing frame is refereed in float64 on the oracle's own latent (``oracle.mimi.rvq_disagreement_margins``); at
    the first differing codebook of a chain the two candidates' distances must agree to ``TIE_TOL`` of |residual|^2, and such
    frames must stay below 0.5 % of all frames.  Returns the number of disagreeing frames."""
    got, ref = got.long().cpu(), ref.long().cpu()
    assert got.shape == ref.shape
    if torch.equal(got, ref):
        return 0
    lat = omimi.encode_latent(clips, mimi_weights)
    margins = omimi.rvq_disagreement_margins(lat, mimi_weights, ref, got)
    assert margins, "codes differ but no chain disagrees?"
    worst = max(m[3] for m in margins)
    assert worst < TIE_TOL, ("not a near-tie", sorted(margins, key=lambda m: -m[3])[:5])
    frames = {(b, f) for b, f, _, _ in margins}
    assert len(frames) <= max(1, int(0.005 * got.shape[0] * got.shape[2])), (len(frames), got.shape)
    return len(frames)
