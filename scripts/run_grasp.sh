#!/bin/bash
# Launch grasp.py from the knobs in scripts/params_script.sh (same contract as the reference launcher:
# empty variables drop their flag, booleans become store_true switches).
source scripts/params_script.sh

flag() { [ -n "$2" ] && echo "$1 $2"; }
switch() { [ "$2" = "true" ] && echo "$1"; }

python grasp.py \
    --model_name_or_path "$MODEL_NAME_OR_PATH" \
    --dataset_name "$DATASET_NAME" \
    --mlp_target_layer_types $MLP_TARGET_LAYER_TYPES \
    --attn_target_layer_types $ATTN_TARGET_LAYER_TYPES \
    --metric "$METRIC" --device "$DEVICE" \
    --num_samples "$NUM_SAMPLES" --batch_size "$BATCH_SIZE" --seq_len "$SEQ_LEN" --padding "$PADDING" \
    --data_path "$DATA_PATH" --train_batch_size "$TRAIN_BATCH_SIZE" --micro_batch_size "$MICRO_BATCH_SIZE" \
    --num_epochs "$NUM_EPOCHS" --learning_rate "$LEARNING_RATE" --max_length "$MAX_LENGTH" \
    --val_set_size "$VAL_SET_SIZE" --prompt_template_name "$PROMPT_TEMPLATE_NAME" \
    --eval_ppl "$EVAL_PPL" --eval_tasks "$EVAL_TASKS" --num_fewshot "$NUM_FEWSHOT" --limit "$LIMIT" \
    --log_file "$LOG_FILE" --train_device "$TRAIN_DEVICE" \
    $(flag --layers_id "$LAYERS_ID") \
    $(flag --num_prune_layers "$NUM_PRUNE_LAYERS") \
    $(flag --compression_ratio "$COMPRESSION_RATIO") \
    $(flag --threshold_ratio "$THRESHOLD_RATIO") \
    $(flag --save_path "$SAVE_PATH") \
    $(flag --resume_from_checkpoint "$RESUME_FROM_CHECKPOINT") \
    $(switch --angular "$ANGULAR") $(switch --allocation_aware "$ALLOCATION_AWARE") \
    $(switch --merge "$MERGE") $(switch --verbose "$VERBOSE") $(switch --recovery "$RECOVERY") \
    $(switch --train_on_inputs "$TRAIN_ON_INPUTS") $(switch --add_eos_token "$ADD_EOS_TOKEN") \
    $(switch --evaluate "$EVALUATE")
