"""CLI driver: ``python -m sesa_audio_separation_b200.inference`` — the flag surface of the reference's
``inference.py`` (proc_folder, :148-236) / ``inference_pytorch.py`` (proc_folder_pytorch_optimized, :277-386) and the
per-file flow of run_folder_pytorch_optimized (:189-274): load -> normalize -> demix -> TTA / demud / instrumental ->
denormalize -> write ``{shortened name}_{instrument}.{wav|flac}``.  The GUI's stdout protocol (``[SESA_PROGRESS]n``)
is kept by demix_pytorch_optimized.  Everything numeric runs on the sm_100a engine; there is no CPU path.
"""
import argparse
import glob
import os
import sys
import time

import numpy as np
import torch

from . import _lib
from .audio_io import load_audio, write_audio
from .backend import create_inference_session
from .config import get_model_from_config, prefer_target_instrument
from .demix import apply_tta, demix, demix_pytorch_optimized, demix_tta, denormalize_audio, normalize_audio


def shorten_filename(filename, max_length=30):
    """inference.py:42-48."""
    base, ext = os.path.splitext(filename)
    if len(base) <= max_length:
        return filename
    return base[:15] + "..." + base[-10:] + ext


def get_soundfile_subtype(pcm_type, is_float=False):
    """inference_pytorch.py:43-52."""
    if is_float:
        return 'FLOAT'
    return {'PCM_16': 'PCM_16', 'PCM_24': 'PCM_24', 'FLOAT': 'FLOAT'}.get(pcm_type, 'FLOAT')


def build_parser():
    p = argparse.ArgumentParser(description="B200 inference for music source separation (SESA CLI surface)")
    p.add_argument("--model_type", type=str, default='mdx23c')
    p.add_argument("--config_path", type=str)
    p.add_argument("--start_check_point", type=str, default='')
    p.add_argument("--input_folder", type=str)
    p.add_argument("--store_dir", type=str, default="")
    p.add_argument("--device_ids", nargs='+', type=int, default=0)
    p.add_argument("--extract_instrumental", action='store_true')
    p.add_argument("--disable_detailed_pbar", action='store_true')
    p.add_argument("--force_cpu", action='store_true')
    p.add_argument("--flac_file", action='store_true')
    p.add_argument("--export_format", type=str, choices=['wav FLOAT', 'flac PCM_16', 'flac PCM_24'], default='flac PCM_24')
    p.add_argument("--pcm_type", type=str, choices=['PCM_16', 'PCM_24'], default='PCM_24')
    p.add_argument("--chunk_size", type=int, default=1000000)   # parsed, never read: chunking comes from the YAML
    p.add_argument("--overlap", type=int, default=4)            # (the reference behaves the same, inference.py:217-233)
    p.add_argument("--optimize_mode", type=str, choices=['channels_last', 'compile', 'jit', 'default'], default='channels_last')
    p.add_argument("--enable_amp", action='store_true')
    p.add_argument("--enable_tf32", action='store_true')
    p.add_argument("--enable_cudnn_benchmark", action='store_true')
    p.add_argument("--lora_checkpoint", type=str, default='')
    p.add_argument("--use_tta", action='store_true')
    p.add_argument("--demud_phaseremix_inst", action='store_true')
    return p


def load_checkpoint_into(model, path, device):
    """inference_pytorch.py:326-369: accept {'state_dict'|'model'|'state': ...} or a bare state_dict, strict=False."""
    try:
        checkpoint = torch.load(path, map_location='cpu', weights_only=True)
    except Exception as e:
        # community checkpoints sometimes pickle arbitrary objects; unpickling those executes code, so it is opt-in
        if os.environ.get('SESA_ALLOW_UNSAFE_CHECKPOINT') != '1':
            raise RuntimeError(f'{path} is not a plain tensor checkpoint ({type(e).__name__}: {e}); set '
                               'SESA_ALLOW_UNSAFE_CHECKPOINT=1 to unpickle it like the reference does '
                               '(this runs code stored in the file)')
        checkpoint = torch.load(path, map_location='cpu', weights_only=False)
    if isinstance(checkpoint, dict):
        for key in ('state_dict', 'model', 'state'):
            if key in checkpoint:
                checkpoint = checkpoint[key]
                break
    missing, unexpected = model.load_state_dict(checkpoint, strict=False)
    matched = len(model.state_dict()) - len(missing)
    print(f"Checkpoint keys: {matched} loaded, {len(missing)} missing, {len(unexpected)} unexpected")
    if matched == 0:
        raise RuntimeError(f'no key of {path} matches the {type(model).__name__} state_dict layout '
                           '(wrong --model_type or config for this checkpoint?)')


def _phase_remix(config, model, args, device, mix_orig, estimates, instruments):
    """"DemudPhaseRemix" instrumental (inference_pytorch.py:233-250): separate a second time from a mix in which the
    lead stem's polarity is flipped (mix -/+ 2*lead) and recombine, so that lead residue cancels instead of adding."""
    lead = 'vocals' if 'vocals' in instruments else instruments[0]
    has_instrumental = 'instrumental' in instruments or 'Instrumental' in instruments

    def second_pass(flipped_mix, tta_base):
        est = demix(config, model, flipped_mix, device, model_type=args.model_type)
        if args.use_tta:
            est = apply_tta(config, model, flipped_mix, tta_base if tta_base is not None else est, device, args.model_type)
        return est

    if not has_instrumental:
        flipped = mix_orig - 2 * estimates[lead]
        return mix_orig + second_pass(flipped, None)[lead]
    flipped = 2 * estimates[lead] - mix_orig
    kept = flipped.copy()
    # (the reference passes the FIRST pass's estimates as the TTA accumulator in this branch, :247)
    return mix_orig + kept - second_pass(flipped, estimates)[lead]


def run_folder(backend, args, config, device, model=None):
    start_time = time.time()
    mixture_paths = sorted(glob.glob(os.path.join(args.input_folder, '*.*')))
    # under torchrun (one process per GPU) the files of the folder are sharded round-robin over the ranks: tracks are
    # independent, so this needs no communication (BASELINE configs 4-5: "track-sharded across 8 x B200")
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    if world > 1:
        mixture_paths = mixture_paths[rank::world]
    sample_rate = getattr(config.audio, 'sample_rate', 44100)
    print(f"B200 backend | {len(mixture_paths)} files | SR: {sample_rate}" + (f" | rank {rank}/{world}" if world > 1 else ""))
    instruments = prefer_target_instrument(config)[:]
    os.makedirs(args.store_dir, exist_ok=True)
    detailed_pbar = not args.disable_detailed_pbar
    written = []
    for path in mixture_paths:
        try:
            mix, sr = load_audio(path, sample_rate)
            if mix.ndim == 1:
                mix = np.stack([mix, mix], axis=0) if int(getattr(config.audio, 'num_channels', 2)) == 2 else mix[None]
            print(f"Loaded audio: {path}, shape: {mix.shape}")
        except Exception as e:
            print(f"Cannot read track: {path}")
            print(f"Error message: {e}")
            continue
        mix_orig = mix.copy()
        norm_params = None
        if 'normalize' in config.inference and config.inference['normalize'] is True:
            mix, norm_params = normalize_audio(mix)
        if args.use_tta and model is not None:
            # demix + apply_tta (:226-229) as one engine run over the three mixes; same stdout protocol
            waveforms_orig = demix_tta(config, model, mix, device, args.model_type,
                                       progress=lambda p: print(f"[SESA_PROGRESS]{p}", flush=True))
            print("[SESA_PROGRESS]100", flush=True)
        else:
            waveforms_orig = demix_pytorch_optimized(config, backend, mix, device, pbar=detailed_pbar)
        if args.demud_phaseremix_inst and model is not None:
            instruments.append('instrumental_phaseremix')
            waveforms_orig['instrumental_phaseremix'] = _phase_remix(
                config, model, args, device, mix_orig, waveforms_orig, instruments)
        if args.extract_instrumental:
            instr = 'vocals' if 'vocals' in instruments else instruments[0]
            waveforms_orig['instrumental'] = mix_orig - waveforms_orig[instr]
            if 'instrumental' not in instruments:
                instruments.append('instrumental')
        for instr in instruments:
            estimates = waveforms_orig[instr]
            if norm_params is not None:
                estimates = denormalize_audio(estimates, norm_params)
            is_float = getattr(args, 'export_format', '').startswith('wav FLOAT')
            codec = 'flac' if getattr(args, 'flac_file', False) else 'wav'
            subtype = get_soundfile_subtype(args.pcm_type, is_float) if codec == 'flac' else get_soundfile_subtype('FLOAT', is_float)
            output_path = os.path.join(args.store_dir, f"{shorten_filename(os.path.basename(path))}_{instr}.{codec}")
            write_audio(output_path, estimates.T, sr, subtype=subtype)
            written.append(output_path)
    print(f"Elapsed time: {time.time() - start_time:.2f} sec")
    return written


def proc_folder(argv=None):
    args = build_parser().parse_args(argv)
    if args.force_cpu:
        raise _lib.SesaError('--force_cpu: the B200 engine has no CPU path')
    _lib.require_cuda()
    if int(os.environ.get('WORLD_SIZE', 1)) > 1 and 'LOCAL_RANK' in os.environ:
        device = f"cuda:{int(os.environ['LOCAL_RANK'])}"          # torchrun: one process per GPU
    else:
        device = f'cuda:{args.device_ids[0]}' if isinstance(args.device_ids, list) else f'cuda:{args.device_ids}'
    print(f"Using device: {device}")
    t0 = time.time()
    model, config = get_model_from_config(args.model_type, args.config_path)
    if args.lora_checkpoint:
        raise NotImplementedError('LoRA checkpoints are out of scope of the inference engine')
    if args.start_check_point != '':
        try:
            load_checkpoint_into(model, args.start_check_point, device)
        except Exception as e:   # corrupt / truncated file: same exit status as the reference (:327-353)
            print(f"CHECKPOINT FILE CORRUPTED\n\nError: {e}\nFile: {args.start_check_point}")
            sys.exit(1)
    model = model.eval().to(device)
    print(f"Instruments: {config.training.instruments}")
    backend = create_inference_session(model=model, device=device, optimize_mode=args.optimize_mode if args.optimize_mode in
                                       ('channels_last', 'default') else 'default', enable_amp=args.enable_amp,
                                       enable_tf32=args.enable_tf32, enable_cudnn_benchmark=args.enable_cudnn_benchmark)
    print(f"Model load time: {time.time() - t0:.2f} sec")
    return run_folder(backend, args, config, device, model=model)


if __name__ == "__main__":
    proc_folder(None)
