"""Drop-in for the reference's ``models.py``: same class names and signatures, B200 CUDA path underneath.
See s-cgib_b200/models.py (hot-path classes) - the other encoders / heads of the reference are out of scope
(SURVEY.md section 8)."""
from scgib_b200.models import GIN, MLP, GINConv, Mainmodel, Mainmodel_continue, Mainmodel_domainadapt, Mainmodel_finetuning  # noqa: F401
