"""Experiment glue around the solve step: the reference's ``src/tools`` entry points that sit directly on either side
of the hot path -- task/result carriers of ``create_data.py`` and the rule-of-thumb rank."""
