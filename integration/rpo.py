"""Drop-in for <RPO checkout>/trainers/rpo.py: copy this file over the reference's trainers/rpo.py
(or put it earlier on sys.path as trainers/rpo.py) with this repository importable.  `train.py`
imports `trainers.rpo` for its side effect -- registering the `RPO` trainer with Dassl
(train.py:31) -- and that is all this module does; the model behind it is rpo_b200's CUDA path.
See INTEGRATION.md."""
from rpo_b200.trainer import RPO, CustomCLIP, PromptLearner, load_clip_to_cpu  # noqa: F401
