"""Drop-in for the reference's ``python/lucas_kanade_reference.py`` (single-scale CLI).

No arithmetic of its own: loads a raw uint8 frame pair, calls the GPU-backed
``lucas_kanade_core`` functions, prints the same kind of statistics, writes
``flow_u.bin`` / ``flow_v.bin`` and the ``x y u v`` text file that
``scripts/visualize_flow.py`` of the reference reads (lucas_kanade_reference.py:78-103).
"""

import argparse
from pathlib import Path
from typing import Optional

import numpy as np
import numpy.typing as npt

from lucas_kanade_core import compute_gradients, lucas_kanade_single_scale

SCRIPT_DIR = Path(__file__).resolve().parent
PROJECT_ROOT = SCRIPT_DIR.parent
DEFAULT_FRAME_DIR = PROJECT_ROOT / "tb" / "test_frames"
DEFAULT_OUTPUT_DIR = SCRIPT_DIR / "output"


def visualize_flow(
    u: npt.NDArray[np.float32], v: npt.NDArray[np.float32], output_path: Path, scale: float = 10.0
) -> None:
    """Quiver plot of every 10th flow vector, coloured by magnitude (needs matplotlib)."""
    import matplotlib.pyplot as plt

    h, w = u.shape
    step = 10
    ys, xs = np.mgrid[step:h:step, step:w:step]
    us, vs = u[step:h:step, step:w:step], v[step:h:step, step:w:step]
    fig, ax = plt.subplots(figsize=(12, 9))
    ax.quiver(xs, ys, us, vs, np.hypot(us, vs), angles="xy", scale_units="xy", scale=1.0 / scale, cmap="jet", width=0.003)
    ax.set_aspect("equal")
    ax.set_xlim(0, w)
    ax.set_ylim(h, 0)
    ax.set_title("Optical Flow Vectors (Single-Scale Lucas-Kanade)")
    ax.set_xlabel("X (pixels)")
    ax.set_ylabel("Y (pixels)")
    plt.colorbar(ax.collections[0], ax=ax, label="Flow Magnitude (pixels)")
    plt.tight_layout()
    plt.savefig(output_path, dpi=150)
    print(f"Flow visualization saved: {output_path}")


def export_flow_field_txt(
    u: np.ndarray,
    v: np.ndarray,
    output_path: Path,
    width: int,
    height: int,
    test_region: Optional[dict] = None,
) -> None:
    """Write '<x> <y> <u> <v>' per pixel (row-major, 6 decimals) after the '#' header lines
    the reference's visualiser expects."""
    header = [
        "# Optical flow field data (Python reference)",
        "# Format: x y u v",
        f"# Image size: {width}x{height}",
    ]
    if test_region:
        header.append(
            f"# Test region: x[{test_region['x_min']}:{test_region['x_max']}], "
            f"y[{test_region['y_min']}:{test_region['y_max']}]"
        )
    uu = np.asarray(u)[:height, :width]
    vv = np.asarray(v)[:height, :width]
    ys, xs = np.mgrid[0:height, 0:width]
    rows = zip(xs.ravel().tolist(), ys.ravel().tolist(), uu.ravel().tolist(), vv.ravel().tolist())
    with open(output_path, "w") as f:
        f.write("\n".join(header) + "\n")
        f.writelines(f"{x} {y} {a:.6f} {b:.6f}\n" for x, y, a, b in rows)
    print(f"Flow field text export: {output_path}")


def main() -> None:
    """CLI with the reference's flags."""
    ap = argparse.ArgumentParser(description="Lucas-Kanade single-scale (B200 backend)")
    ap.add_argument("--frame-dir", type=str, default=str(DEFAULT_FRAME_DIR), help="Directory containing frame_00.bin and frame_01.bin")
    ap.add_argument("--width", type=int, default=320, help="Frame width")
    ap.add_argument("--height", type=int, default=240, help="Frame height")
    ap.add_argument("--window-size", type=int, default=5, help="Window size for Lucas-Kanade")
    ap.add_argument("--output-dir", type=str, default=str(DEFAULT_OUTPUT_DIR), help="Output directory for results")
    args = ap.parse_args()

    out_dir = Path(args.output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    frame_dir = Path(args.frame_dir)
    shape = (args.height, args.width)
    frame_prev = np.fromfile(frame_dir / "frame_00.bin", dtype=np.uint8).reshape(shape).astype(np.float32)
    frame_curr = np.fromfile(frame_dir / "frame_01.bin", dtype=np.uint8).reshape(shape).astype(np.float32)
    print(f"Loaded frames: {args.width}x{args.height}")
    print(f"Window size: {args.window_size}x{args.window_size}")

    print("\nComputing gradients...")
    Ix, Iy, It = compute_gradients(frame_prev, frame_curr)
    print("\nGradient statistics:")
    for name, g in (("Ix", Ix), ("Iy", Iy), ("It", It)):
        print(f"  {name} range: [{np.min(g):.2f}, {np.max(g):.2f}]")

    print("Computing optical flow...")
    u, v = lucas_kanade_single_scale(frame_prev, frame_curr, window_size=args.window_size)
    half = args.window_size // 2
    interior = u[half : args.height - half, half : args.width - half]
    print("\nWindow analysis:")
    print(f"  Total possible windows: {interior.size}")
    print(f"  Windows with non-zero flow: {int(np.count_nonzero(interior))}")

    region = np.s_[105:135, 55:85]
    print("\n=== Results ===")
    print(f"Mean flow in square region: u={np.mean(u[region]):.3f}, v={np.mean(v[region]):.3f}")
    print(f"Std dev in square region:   u={np.std(u[region]):.3f}, v={np.std(v[region]):.3f}")
    print("Expected: u=2.0, v=0.0")  # the reference prints this whatever the frames are (lucas_kanade_reference.py:176)

    u.tofile(out_dir / "flow_u.bin")
    v.tofile(out_dir / "flow_v.bin")
    print(f"\nFlow fields saved to {out_dir}")
    export_flow_field_txt(
        u=u, v=v, output_path=out_dir / "flow_field_python.txt", width=args.width, height=args.height,
        test_region={"x_min": 55, "x_max": 85, "y_min": 105, "y_max": 135},
    )
    try:
        visualize_flow(u, v, out_dir / "flow_visualization_single_scale.png")
    except ImportError:
        print("Matplotlib not available, skipping visualization")


if __name__ == "__main__":
    main()
