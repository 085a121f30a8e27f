"""Reference arm of ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` leg): the UNMODIFIED reference
(HaoIrving/RefineDet.PyTorch) run on the host cores.  ``build_ref.py`` copies the files of the path from the
reference checkout into the git-ignored ``baseline/_ref/`` (it travels to the GPU box like ``oracle/_ref/``);
``reference_arm.py`` drives them.  Nothing under ``refinedet/`` imports this package."""
