"""python -m compressai.utils.eval_model -a stf -p checkpoint.pth.tar -d images/ [--entropy-estimation] [-r recon/]"""
import argparse
import json
import sys

from . import collect_images, eval_model, load_checkpoint


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("-d", "--dataset", required=True, help="directory of images")
    ap.add_argument("-a", "--architecture", default="stf", help="zoo model: stf, cnn, cnn2")
    ap.add_argument("-p", "--path", required=True, help="checkpoint (.pth.tar with a 'state_dict' entry)")
    ap.add_argument("-r", "--recon_path", default=None, help="where to save reconstructions")
    ap.add_argument("--entropy-estimation", action="store_true", help="forward() and likelihood-based bpp instead of real coding")
    args = ap.parse_args(argv)
    files = collect_images(args.dataset)
    if not files:
        print("no images found", file=sys.stderr)
        return 1
    model = load_checkpoint(args.architecture, args.path)
    res = eval_model(model, files, entropy_estimation=args.entropy_estimation, recon_path=args.recon_path)
    print(json.dumps({"name": args.architecture, "description": "Inference (entropy estimation)" if args.entropy_estimation else "Inference (ans)",
                      "results": res}, indent=2))
    return 0


if __name__ == "__main__":
    sys.exit(main())
