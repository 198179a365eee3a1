"""Copy the shipped checkpoints the parity tests use from /root/reference into
checkpoints/_ref/ (git-ignored data, not source; travels to the GPU box with the snapshot).
Build container only."""
from __future__ import annotations

import os
import shutil

from . import ref_shims

# Every checkpoint a parity test or bench.py loads.  (The CRN and LSTM files are 70 / 87 MB; bench.py refuses to
# run without the CRN one.)
WANTED = [("CRN", "wsj0_si84_300h_crn_noncprs_model.pth"), ("LSTM", "vb_lstm_noncprs_model.pth"),
          ("DCCRN_SNR", "wsj0_si84_300h_dccrn_snr_model.pth"),
          ("FullSubNet", "wsj0_si84_300h_fullsubnet_cprs_model_512_256.pth"),
          ("DCCRN", "wsj0_si84_300h_dccrn_cprs_model.pth"),
          ("Uformer", "wsj0_si84_300h_uformer_noncprs_model.pth"),
          ("GCRN", "vb_gcrn_cprs_model.pth"), ("DPCRN", "vb_dpcrn_noncprs_model.pth"),
          ("CTSNet", "step1_vb_cts_noncprs_model_final.pth"), ("CTSNet", "step2_vb_cts_noncprs_model.pth"),
          ("CTSNet_new", "step1_vb_cts_cprs_model_final.pth"), ("CTSNet_new", "step2_vb_cts_cprs_model.pth"),
          ("TaylorSENet", "vb_taylor_noncprs_model.pth"), ("TaylorSENet_new", "vb_taylor_cprs_model.pth"),
          ("G2Net_new", "vb_gaf_cprs_model.pth"), ("G2Net_VB", "vb_gaf_noncprs_model.pth")]
DEST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "checkpoints", "_ref")


def main():
    os.makedirs(DEST, exist_ok=True)
    for mdir, name in WANTED:
        src = ref_shims.checkpoint_path(mdir, name)
        dst = os.path.join(DEST, f"{mdir}__{name}")
        if not os.path.exists(dst):
            shutil.copyfile(src, dst)
        print(dst, os.path.getsize(dst))


def path_for(mdir: str, name: str) -> str:
    return os.path.join(DEST, f"{mdir}__{name}")


if __name__ == "__main__":
    main()
