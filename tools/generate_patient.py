"""One synthetic patient, command-line compatible with the reference's
synthetic_gens/generatePatient_FK_2c.py (:160-187): loads the tissue NIfTI files, runs the FK_2c
solver on the GPU, writes P / N / S / wm_data / gm_data / segm .nii.gz and parameters.json / .txt.
The FET-PET surrogate (a sequential Gibbs noise sweep, :10-53) is out of scope and not written.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser(description='Generate Synthetic Patient Data (B200)')
    for name in ('th_necro_n', 'th_enhancing_p', 'th_edema_p', 'Dw', 'rho', 'lambda_np', 'sigma_np', 'D_s', 'lambda_s',
                 'NxT1_pct', 'NyT1_pct', 'NzT1_pct', 'resolution_factor'):
        ap.add_argument('--' + name, type=float, required=True)
    ap.add_argument('--gm_path', type=str, required=True)
    ap.add_argument('--wm_path', type=str, required=True)
    ap.add_argument('--RatioDw_Dg', type=float, default=10.0)
    ap.add_argument('--th_matter', type=float, default=0.1)
    ap.add_argument('--th_necro', type=float, default=0.9)
    ap.add_argument('--dx_mm', type=float, default=1.0)
    ap.add_argument('--dy_mm', type=float, default=1.0)
    ap.add_argument('--dz_mm', type=float, default=1.0)
    ap.add_argument('--init_scale', type=float, default=1.0)
    ap.add_argument('--time_series_solution_Nt', type=int)
    ap.add_argument('--Nt_multiplier', type=int, default=10)
    ap.add_argument('--output_dir', type=str, required=True)
    ap.add_argument('--precision', type=str, default='fp64')
    ap.add_argument('--device', type=int, default=0)
    args = ap.parse_args()

    from tumorgrowthtoolkit_b200 import nifti
    from tumorgrowthtoolkit_b200.FK_2c import Solver
    from tumorgrowthtoolkit_b200.ensemble import create_segmentation_map
    os.makedirs(args.output_dir, exist_ok=True)
    args_dict = vars(args)
    with open(os.path.join(args.output_dir, 'parameters.json'), 'w') as f:
        json.dump(args_dict, f, indent=4)
    with open(os.path.join(args.output_dir, 'parameters.txt'), 'w') as f:
        for k, v in args_dict.items():
            f.write(f"{k}: {v}\n")
    gm, wm = nifti.load(args.gm_path), nifti.load(args.wm_path)
    keys = ('Dw', 'rho', 'lambda_np', 'sigma_np', 'D_s', 'lambda_s', 'NxT1_pct', 'NyT1_pct', 'NzT1_pct', 'resolution_factor',
            'RatioDw_Dg', 'th_matter', 'th_necro', 'dx_mm', 'dy_mm', 'dz_mm', 'init_scale', 'time_series_solution_Nt',
            'Nt_multiplier', 'precision', 'device')
    params = {k: args_dict[k] for k in keys}
    params.update(gm=gm, wm=wm)
    result = Solver(params).solve()
    fs = result['final_state']
    for k in 'PNS':
        nifti.save(fs[k], os.path.join(args.output_dir, f'{k}.nii.gz'))
    nifti.save(wm, os.path.join(args.output_dir, 'wm_data.nii.gz'))
    nifti.save(gm, os.path.join(args.output_dir, 'gm_data.nii.gz'))
    seg = create_segmentation_map(fs['P'], fs['N'], args.th_necro_n, args.th_enhancing_p, args.th_edema_p)
    nifti.save(seg, os.path.join(args.output_dir, 'segm.nii.gz'))
    print(f"Patient data saved in {args.output_dir}")


if __name__ == '__main__':
    main()
