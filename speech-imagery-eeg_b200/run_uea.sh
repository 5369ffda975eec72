#!/bin/bash
# Same driver as the reference's run_uea.sh: 30 UEA archives x 5 seeds, InterpGN(FCN).
# Without the archives on disk, run.py draws synthetic series of each archive's shape.
# Multi-GPU: NPROC=8 bash run_uea.sh   (one process per GPU, NCCL gradient all-reduce)
MODEL=InterpGN
DNN_TYPE=FCN
NUM_SHAPELET=10
LAMBDA_DIV=0.1
LAMBDA_REG=0.1
EPS=1
BETA_SCHEDULE=constant
GATING_VALUE=1
EPOCHS=${EPOCHS:-500}
NPROC=${NPROC:-1}

UEA_DATASETS=(
    ArticularyWordRecognition AtrialFibrillation BasicMotions CharacterTrajectories LSST ERing Epilepsy
    EthanolConcentration FaceDetection FingerMovements Handwriting Heartbeat InsectWingbeat JapaneseVowels
    Libras NATOPS PenDigits RacketSports SpokenArabicDigits UWaveGestureLibrary Cricket PhonemeSpectra
    HandMovementDirection SelfRegulationSCP1 SelfRegulationSCP2 StandWalkJump PEMS-SF DuckDuckGeese
    MotorImagery EigenWorms
)
SEEDS=(0 42 1234 8237 2023)

if [ "$NPROC" -gt 1 ]; then
    LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NPROC --master-addr 127.0.0.1 --master-port ${PORT:-29511}"
else
    LAUNCH="python"
fi

cd "$(dirname "$0")"
for dataset in ${DATASETS:-${UEA_DATASETS[@]}}; do
    for seed in ${SEEDS[@]}; do
        $LAUNCH run.py \
            --model $MODEL --dnn_type $DNN_TYPE --dataset $dataset --train_epochs $EPOCHS --batch_size 32 \
            --lr 5e-3 --dropout 0. --num_shapelet $NUM_SHAPELET --lambda_div $LAMBDA_DIV --lambda_reg $LAMBDA_REG \
            --epsilon $EPS --beta_schedule $BETA_SCHEDULE --seed $seed --gating_value $GATING_VALUE --amp
    done
done
