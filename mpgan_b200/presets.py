"""Model factories with the reference's default wiring.

``mp_generator`` / ``mp_discriminator`` build exactly what ``setup_training.setup_mpgan``
(setup_training.py:1195-1293) builds from ``trained_models/mp_g/args.txt``; ``gapt_generator`` /
``gapt_discriminator`` mirror ``setup_gapt`` (:1296-1347) with its argparse defaults (:551-617).
Keyword overrides use the reference's ``args`` names.
"""
from __future__ import annotations

from .gapt import GAPT_D, GAPT_G
from .model import MPDiscriminator, MPGenerator

MP_DEFAULTS = dict(
    leaky_relu_alpha=0.2, gen_dropout=0.0, disc_dropout=0.5, batch_norm_gen=False, batch_norm_disc=False,
    spectral_norm_gen=False, spectral_norm_disc=False, pos_diffs=False, all_ef=False, coords="polarrel",
    deltacoords=False, deltar=False, int_diffs=False, clabels=0, mask_fne_np=False, fully_connected=True,
    num_knn=10, self_loops=True, sum=True, clabels_first_layer=0, num_hits=30, hidden_node_size=32,
    fe=[96, 160, 192], fn=[256, 256], mp_iters_gen=2, mp_iters_disc=2, fe1g=0, fe1d=0, gtanh=True,
    node_feat_size=3, latent_node_size=32, lfc=False, lfc_latent_size=128, loss="ls", dea=True, fnd=[],
    mask_fnd_np=False, mask_feat=False, mask_feat_bin=False, mask_weights=False, mask_manual=False,
    mask_exp=False, mask_real_only=False, mask_learn=False, mask_learn_bin=True, mask_learn_sep=False,
    fmg=[64], mask_disc_sep=False, mask_c=True,
)

GAPT_DEFAULTS = dict(
    leaky_relu_alpha=0.2, gen_dropout=0.0, disc_dropout=0.5, batch_norm_gen=False, batch_norm_disc=False,
    spectral_norm_gen=False, spectral_norm_disc=False, num_hits=30, num_heads=4, gapt_embed_dim=64,
    sab_fc_layers=[], gapt_mask=True, use_isab=False, num_isab_nodes=10, sab_layers_gen=4, sab_layers_disc=2,
    node_feat_size=3, final_fc_layers_gen=[], final_fc_layers_disc=[], layer_norm_gen=False,
    layer_norm_disc=False,
)


def _mp_parts(a, gen):
    linear_args = {
        "leaky_relu_alpha": a["leaky_relu_alpha"],
        "dropout_p": a["gen_dropout"] if gen else a["disc_dropout"],
        "batch_norm": a["batch_norm_gen"] if gen else a["batch_norm_disc"],
        "spectral_norm": a["spectral_norm_gen"] if gen else a["spectral_norm_disc"],
    }
    mp_args = {
        "pos_diffs": a["pos_diffs"], "all_ef": a["all_ef"], "coords": a["coords"],
        "delta_coords": a["deltacoords"], "delta_r": a["deltar"], "int_diffs": a["int_diffs"],
        "clabels": a["clabels"], "mask_fne_np": a["mask_fne_np"], "fully_connected": a["fully_connected"],
        "num_knn": a["num_knn"], "self_loops": a["self_loops"], "sum": a["sum"],
    }
    common = {
        "num_particles": a["num_hits"], "hidden_node_size": a["hidden_node_size"], "fe_layers": a["fe"],
        "fn_layers": a["fn"], "fn1_layers": None,
    }
    mask_args = {k: a[k] for k in (
        "mask_feat", "mask_feat_bin", "mask_weights", "mask_manual", "mask_exp", "mask_real_only", "mask_learn",
        "mask_learn_bin", "mask_learn_sep", "fmg", "mask_disc_sep", "mask_fnd_np", "mask_c", "mask_fne_np")}
    return linear_args, mp_args, common, mask_args


def mp_generator(**over) -> MPGenerator:
    a = {**MP_DEFAULTS, **over}
    linear_args, mp_args, common, mask_args = _mp_parts(a, True)
    return MPGenerator(
        mp_iters=a["mp_iters_gen"], fe1_layers=a["fe1g"] if a["fe1g"] else None,
        final_activation="tanh" if a["gtanh"] else "", output_node_size=a["node_feat_size"],
        input_node_size=a["latent_node_size"], lfc=a["lfc"], lfc_latent_size=a["lfc_latent_size"],
        **common, mp_args=mp_args, mp_args_first_layer={"clabels": a["clabels_first_layer"]},
        linear_args=linear_args, mask_args=mask_args)


def mp_discriminator(**over) -> MPDiscriminator:
    a = {**MP_DEFAULTS, **over}
    linear_args, mp_args, common, mask_args = _mp_parts(a, False)
    return MPDiscriminator(
        mp_iters=a["mp_iters_disc"], fe1_layers=a["fe1d"] if a["fe1d"] else None,
        final_activation="" if a["loss"] in ("w", "hinge") else "sigmoid", input_node_size=a["node_feat_size"],
        dea=a["dea"], dea_sum=a["sum"], fnd=a["fnd"], mask_fnd_np=a["mask_fnd_np"],
        **common, mp_args=mp_args,
        mp_args_first_layer={"clabels": a["clabels_first_layer"], "all_ef": False},
        linear_args=linear_args, mask_args=mask_args)


def _gapt_parts(a, gen):
    linear_args = {
        "leaky_relu_alpha": a["leaky_relu_alpha"],
        "dropout_p": a["gen_dropout"] if gen else a["disc_dropout"],
        "batch_norm": a["batch_norm_gen"] if gen else a["batch_norm_disc"],
        "spectral_norm": a["spectral_norm_gen"] if gen else a["spectral_norm_disc"],
    }
    common = {
        "num_particles": a["num_hits"], "num_heads": a["num_heads"], "embed_dim": a["gapt_embed_dim"],
        "sab_fc_layers": a["sab_fc_layers"], "use_mask": a["gapt_mask"], "use_isab": a["use_isab"],
        "num_isab_nodes": a["num_isab_nodes"],
    }
    return linear_args, common


def gapt_generator(**over) -> GAPT_G:
    a = {**GAPT_DEFAULTS, **over}
    linear_args, common = _gapt_parts(a, True)
    return GAPT_G(sab_layers=a["sab_layers_gen"], output_feat_size=a["node_feat_size"],
                  final_fc_layers=a["final_fc_layers_gen"], dropout_p=a["gen_dropout"],
                  layer_norm=a["layer_norm_gen"], **common, linear_args=linear_args)


def gapt_discriminator(**over) -> GAPT_D:
    a = {**GAPT_DEFAULTS, **over}
    linear_args, common = _gapt_parts(a, False)
    return GAPT_D(sab_layers=a["sab_layers_disc"], input_feat_size=a["node_feat_size"],
                  final_fc_layers=a["final_fc_layers_disc"], dropout_p=a["disc_dropout"],
                  layer_norm=a["layer_norm_disc"], **common, linear_args=linear_args)
