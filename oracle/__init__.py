"""CPU oracle for the DVG stochastic-rollout hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``dvg_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker or as
the timed CPU baseline -- never as the product path.

Pieces (each cites the reference file:line it restates):

* ``lstm_ref``    -- ``models/lstm.py:42-72`` (``lstm``) and ``:140-175``
                     (``gaussian_lstm``).  PINNED: checked against golden
                     vectors produced by importing the real reference module
                     (``tests/golden/make_golden.py``).
* ``gp_ref``      -- ``models/gp_models.py:10-24`` + the gpytorch 0.3.x
                     ``WhitenedVariationalStrategy`` / ``GaussianLikelihood``
                     eval-mode predictive it calls into.  gpytorch is a
                     third-party dependency that is absent from the reference
                     tree and from this image (no version pin in the reference;
                     the API used exists only in gpytorch >=0.3.0,<1.0), so this
                     part is a restatement of the published algorithm:
                     **PARITY UNPINNED** for the GP stage.
* ``trigger_ref`` -- ``generate_frames.py:227-232,275,283-289`` (numpy fp32).
* ``rollout_ref`` -- ``generate_frames.py:107-178`` (make_gifs),
                     ``:249-300`` (GPtrigger_gen), ``train.py:256-289`` (plot)
                     with injectable noise.
"""
