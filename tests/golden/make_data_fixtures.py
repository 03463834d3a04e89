"""Extract the on-disk inputs of the reference's real-data configurations into small fixtures (run in the build
container, where /root/reference exists; the GPU box only sees the .npz files):

    python tests/golden/make_data_fixtures.py

  data_mouseob.npz   C3  MouseOB spatial: 455 genes of data/MouseOB/SE_genes_comparison.csv whose counts
                         (mouse_ob_SV_genes.csv), scale factors (scale_nb_model_sel.csv) and coordinates (locations.csv)
                         are on disk, with the LLR / p / q the reference published for them (LLR_gpcounts ...).
  data_mousehsc.npz  C2 / C4  MouseHSC: 19 genes x 2660 cells, Slingshot pseudotime and curve weights
                         (Branching_GPcounts.ipynb cells 2-5: label 1 where curve1 > 0.8 else 2, every 5th cell from 1).
  data_fission.npz   C1  fission yeast bulk time course: 24 seeded-random genes x 36 samples, minutes
                         (bulk_time_series.ipynb: One_sample_test on the first 18 columns, Two_samples_test on all 36).
"""
import os

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/data/"


def mouseob():
    loc = pd.read_csv(REF + "MouseOB/locations.csv", index_col=0)
    Yall = pd.read_csv(REF + "MouseOB/mouse_ob_SV_genes.csv", index_col=0)
    Sall = pd.read_csv(REF + "MouseOB/scale_nb_model_sel.csv", index_col=0)
    paper = pd.read_csv(REF + "MouseOB/SE_genes_comparison.csv", index_col=0).set_index("gene")
    genes = [g for g in paper.index if g in Yall.columns and g in Sall.columns]
    X = loc.iloc[:, :2].values.astype(np.float64)
    Y = Yall[genes].values.T
    assert np.all(Y == np.round(Y))
    scale = Sall[genes].values.astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "data_mouseob.npz"), X=X, Y=Y.astype(np.int32), scale=scale,
                        genes=np.array(genes), paper_llr=paper.loc[genes, "LLR_gpcounts"].values,
                        paper_p=paper.loc[genes, "pval_gpcounts"].values, paper_q=paper.loc[genes, "qval_gpcounts"].values)
    print("mouseob", X.shape, Y.shape, scale.shape, len(genes))


def mousehsc():
    data = pd.read_csv(REF + "MouseHSC/geneExpression.csv", index_col=[0])
    sl = pd.read_csv(REF + "MouseHSC/Slingshot.csv", index_col=[0])
    assert list(data.columns) == list(sl.index)
    np.savez_compressed(os.path.join(HERE, "data_mousehsc.npz"), Y=data.values.astype(np.float64),
                        genes=np.array(list(data.index)), pseudotime=sl["pseudotime"].values.astype(np.float64),
                        curve1=sl["curve1"].values.astype(np.float64), curve2=sl["curve2"].values.astype(np.float64))
    print("mousehsc", data.shape)


def fission():
    Y = pd.read_csv(REF + "fission_normalized_counts.csv", index_col=[0])
    col = pd.read_csv(REF + "fission_col_data.csv", index_col=[0])
    assert list(Y.columns) == list(col.index)
    rng = np.random.default_rng(2024)
    sel = np.sort(rng.choice(Y.shape[0], 24, replace=False))
    sel = np.r_[sel, [list(Y.index).index("SPAC11D3.01c")]]        # the notebook's gene
    np.savez_compressed(os.path.join(HERE, "data_fission.npz"), Y=Y.values[sel].astype(np.float64),
                        genes=np.array(list(Y.index[sel])), minute=col["minute"].values.astype(np.float64),
                        strain=np.array(list(col["strain"])))
    print("fission", Y.shape, len(sel))


if __name__ == "__main__":
    mouseob()
    mousehsc()
    fission()
