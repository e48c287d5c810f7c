"""The UNMODIFIED reference as a timed / checked baseline (test + bench infrastructure, never product code).

``baseline/_ref/`` (git-ignored, not gpurun-ignored: it travels to the GPU box) holds verbatim copies of the
reference's ``models.py``, ``utils.py``, ``main.py``, ``dataset.py`` and ``config/*.yml``, made by
``python -m baseline.install`` in the build container.  ``baseline.loader`` imports them with the two stubs the
survey documents (torchtext / h5py are absent in this image) and the one-token ``BCELoss(reduction=None)`` reading.
Only ``tests/``, ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs and ``tools/`` import this package.
"""
