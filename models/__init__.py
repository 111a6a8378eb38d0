"""Drop-in ``models`` package: put this repository ahead of the reference checkout on sys.path and
``from models.adamvs import AdaMVSNet, Infer_AdaMVSNet, cas_mvs_vis_loss`` resolves here."""
