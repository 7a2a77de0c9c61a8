"""Reference-compatible entry point: `python main_prior_training.py --train_dir ... --max_bitrate ...`
(flags as in the reference driver, main_prior_training.py:11-21)."""
from recombiner_b200.main_prior_training import main, parse_args, save_checkpoint, train_prior  # noqa: F401

if __name__ == '__main__':
    main()
